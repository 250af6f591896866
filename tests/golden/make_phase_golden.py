"""Generates tests/golden/phase_data_ref.npz from the REFERENCE's own code: the method
SimulationConditionsBASE.CalculatePhaseData is cut out of /root/reference/TranscranialModeling/BabelIntegrationBASE.py
with `ast` (the module itself cannot be imported here: nibabel, h5py, SimpleITK, BabelViscoFDTD are absent) and executed,
unmodified, on a stub object holding seeded inputs.  Nothing of the reference is copied into the repository; only the
inputs and the arrays it produced are stored.

    python tests/golden/make_phase_golden.py          (needs /root/reference; run in the build container only)
"""
import ast
import os
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/TranscranialModeling/BabelIntegrationBASE.py'


def reference_method():
    src = open(REF).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == 'CalculatePhaseData':
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {'np': np, 'fft': np.fft, 'time': time}      # BabelIntegrationBASE.py:10, :34-37, :25
            exec(compile(mod, REF, 'exec'), ns)
            return ns['CalculatePhaseData'], (node.lineno, node.end_lineno)
    raise RuntimeError('CalculatePhaseData not found in the reference')


def make_inputs(seed=7, shape=(9, 7, 11), ppp=48, sub=12, periods=2, frequency=500e3, pml=2):
    rng = np.random.default_rng(seed)
    n1, n2, n3 = shape
    dt = 1.0 / frequency / ppp
    nsamples = periods * ppp // sub
    first = 37                                                   # SensorStart: any whole number of samples
    t = (first + np.arange(nsamples)) * sub * dt
    sensor_map = np.zeros(shape, np.uint32)
    sensor_map[pml:-pml, pml:-pml, pml + 1:-pml] = 1
    sensor_map[rng.random(shape) < 0.15] = 0                     # ragged: holes in the sensor box
    index = (np.flatnonzero(sensor_map.reshape(-1, order='F')) + 1).astype(np.uint32)
    amp = rng.uniform(0.0, 2.0e5, index.size)
    ph = rng.uniform(-np.pi, np.pi, index.size)
    w = 2 * np.pi * frequency
    p = (amp[:, None] * np.sin(w * t[None, :] + ph[:, None]) + 0.05 * amp[:, None] * np.sin(2 * w * t[None, :] + 1.0)
         + rng.normal(0, 50.0, (index.size, nsamples)))
    return dict(shape=np.array(shape), ppp=ppp, sub=sub, frequency=frequency, time=t, pressure=p.astype(np.float32), index=index)


def main():
    fn, lines = reference_method()
    inp = make_inputs()
    n1, n2, n3 = (int(x) for x in inp['shape'])
    me = types.SimpleNamespace(_N1=n1, _N2=n2, _N3=n3, _Sensor={'time': inp['time'].copy(), 'Pressure': inp['pressure'].copy()},
                               _PPP=inp['ppp'], _SensorSubSampling=inp['sub'], _Frequency=inp['frequency'],
                               _InputParam=inp['index'].copy(), _DictPeakValue={'Pressure': None})
    fn(me, bRefocused=False, bDoRefocusing=False)
    out = os.path.join(HERE, 'phase_data_ref.npz')
    np.savez_compressed(out, PhaseMap=me._PhaseMap, PressMapFourier=me._PressMapFourier, PressMapPeak=me._PressMapPeak,
                        reference_lines=np.array(lines), **inp)
    print('wrote', out, 'from', REF, 'lines %d-%d' % lines, 'numpy', np.__version__)


if __name__ == '__main__':
    main()
