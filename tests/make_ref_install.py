"""Places an unmodified copy of the reference's TranscranialModeling and ThermalModeling packages under baseline/_ref/ (git-ignored, so it never
enters the repository history; not gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist).
tests/test_reference_caller.py imports the reference caller from there.  __graft_entry__.build() runs this whenever
/root/reference is present.

    python tests/make_ref_install.py
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference/TranscranialModeling'
DST = os.path.join(ROOT, 'baseline', '_ref', 'TranscranialModeling')


def install():
    if not os.path.isdir(SRC):
        return None
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    # the Python sources of the caller only: geometry tables (.mat/.csv) are not touched by the harness
    # (the two small HDF5 files travel too: MapPichardo.h5 is read at import time, and both pin babelbrain_b200/h5mini.py)
    ignore = shutil.ignore_patterns('__pycache__', '*.mat', '*.csv', '*.stl', '*.npz')
    shutil.copytree(SRC, DST, ignore=ignore)
    thermal_src, thermal_dst = os.path.join(os.path.dirname(SRC), 'ThermalModeling'), os.path.join(os.path.dirname(DST), 'ThermalModeling')
    if os.path.isdir(thermal_src):         # the driver of the thermal step (tests/test_reference_thermal_caller.py)
        if os.path.isdir(thermal_dst):
            shutil.rmtree(thermal_dst)
        shutil.copytree(thermal_src, thermal_dst, ignore=ignore)
    return DST


if __name__ == '__main__':
    print(install() or ('no reference tree at %s' % SRC))
    sys.exit(0)
