"""Generates the golden vectors of tests/golden/*.npz.

PARITY UNPINNED: the reference implementation of this path (the BabelViscoFDTD package) is not in
/root/reference and cannot be installed here, so these vectors come from the float64 NumPy oracle
(oracle/fdtd_numpy.py), not from the reference.  They pin the oracle (and through it the CUDA path)
against drift; when BabelViscoFDTD becomes reachable, rerun with --reference to regenerate them from
`PropagationModel.StaggeredFDTD_3D_with_relaxation(COMPUTING_BACKEND=0)` on the same inputs.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {
    # name: (workload, shape, periods, pml, extra kwargs)
    'water_focus': ('single_water', (24, 22, 30), 3, 4, {}),
    'skull_plane': ('ctx500_skull', (26, 24, 34), 3, 5, {}),
    'skull_allmaps': ('ctx500_skull', (22, 24, 28), 2, 4,
                      dict(SelMapsRMSPeakList=['ALLV', 'Vx', 'Vy', 'Vz', 'Sigmaxx', 'Sigmayy', 'Sigmazz', 'Sigmaxy', 'Sigmaxz', 'Sigmayz', 'Pressure'],
                           SelMapsSensorsList=['Vx', 'Vz', 'Sigmaxx', 'Sigmaxy', 'Pressure'], SelRMSorPeak=3)),
    'dome_stress': ('dome_stress', (28, 28, 24), 2, 4, {}),
}
DROP = ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')


def inputs_digest(w):
    h = hashlib.sha256()
    for a in w['args']:
        h.update(np.ascontiguousarray(np.asarray(a)).tobytes())
    for k in sorted(w['kwargs']):
        v = w['kwargs'][k]
        if isinstance(v, np.ndarray):
            h.update(np.ascontiguousarray(v).tobytes())
        else:
            h.update(repr(v).encode())
    return h.hexdigest()


def build_case(name):
    from babelbrain_b200 import workloads
    wl, shape, periods, pml, extra = CASES[name]
    w = workloads.make_workload(wl, shape=shape, periods=periods, pml=pml)
    w['kwargs'].update(extra)
    return w


def main():
    from oracle import fdtd_numpy
    for name in CASES:
        w = build_case(name)
        kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
        r = fdtd_numpy.run(*w['args'], dtype=np.float64, **kw)
        out = {'digest': np.array(inputs_digest(w)), 'IndexSensorMap': r['IndexSensorMap'], 'steps': np.array(r['steps']),
               'time': r['Sensor']['time']}
        for k, v in r['RMS'].items():
            out['RMS_' + k] = v.astype(np.float32)
        for k, v in r['Peak'].items():
            out['Peak_' + k] = v.astype(np.float32)
        for k, v in r['Sensor'].items():
            if k != 'time':
                out['Sensor_' + k] = v.astype(np.float32)
        for k in ('Vz', 'Sigmaxx', 'Pressure'):
            out['Last_' + k] = r['LastMap'][k].astype(np.float32)
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, 'KiB', 'steps', r['steps'])


if __name__ == '__main__':
    main()
