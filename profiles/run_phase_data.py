"""Phase / amplitude maps of a finished CTX-500 simulation: the caller's host route (download Sensor['Pressure'] and
IndexSensorMap, FFT on the host cores as BabelIntegrationBASE.py:2498-2518 does -- timed with the oracle restatement of
that method) against bb_fdtd_get_phase_data (single-bin DFT on the device, three volumes downloaded).
   python profiles/run_phase_data.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import FdtdSlab, PropagationModel
from oracle import phase_data
w = workloads.make_workload('ctx500_skull')
m = w['meta']
kw = {k: v for k, v in w['kwargs'].items() if k not in ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')}
s = FdtdSlab(*w['args'], **kw)
s.run()
PM = PropagationModel()
PM._last_slabs = (s,)
for rep in range(3):
    t0 = time.perf_counter()
    P = s.get_sensors('Pressure')
    t1 = time.perf_counter()
    ph, fo, pk = phase_data.calculate_phase_data(s.sample_steps * s.dt, P, s.IndexSensorMap, s.shape, m['frequency'], m['ppp'], m['sub'])
    t2 = time.perf_counter()
    res = PM.CalculatePhaseDataOnDevice(m['frequency'])
    t3 = time.perf_counter()
    err = np.abs(res['PressMapFourier'] - fo).max() / np.abs(fo).max()
    print('host route: download %.3f s (%.0f MB) + FFT/scatter %.3f s | device route: %.3f s (%.0f MB) | max |dF|/max|F| = %.1e, peak maps equal: %s'
          % (t1 - t0, (P.nbytes + s.IndexSensorMap.nbytes) / 1e6, t2 - t1, t3 - t2, (fo.nbytes + ph.nbytes + pk.nbytes) / 1e6, err,
             np.array_equal(res['PressMapPeak'], pk)), flush=True)
    del P, ph, fo, pk, res
