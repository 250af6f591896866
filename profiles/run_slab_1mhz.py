"""One rank's share of the 1 MHz high-resolution domain (BASELINE configs[4]: 1080^3 cut into 8 slabs of 135 planes),
run as a stand-alone 135 x 1080 x 1080 domain for a few periods: per-GPU throughput on the north-star shape.
   python profiles/run_slab_1mhz.py [planes] [time_steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import FdtdSlab
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 135
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
t0 = time.time()
w = workloads.make_workload('hires_1mhz', shape=(planes, 1080, 1080), periods=4)
print('built', w['meta']['shape'], 'in %.1f s' % (time.time() - t0), flush=True)
kw = {k: v for k, v in w['kwargs'].items() if k not in ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')}
kw['SensorStart'] = 10 ** 6 // kw['SensorSubSampling']     # keep the RMS window out of the timed steps
MM, ML = w['args'][0], w['args'][1]
cls, alg = bench.traffic_model(MM, ML, 12, 0, planes, 0, planes)
s = FdtdSlab(*w['args'], **kw)
print('device bytes %.2f GB' % (s.stats()['device_bytes'] / 1e9), flush=True)
s.run(3)
st = s.run(n, profile=True)
cells = w['meta']['cells']
peak = bench.measured_peak()[0]
print('classes', {k: round(v / cells, 3) for k, v in cls.items()})
print('per step: stress %.3f ms (%.0f GB/s, %.2f of peak)  particle %.3f ms (%.0f GB/s, %.2f)  total %.3f ms -> %.1f Gcell-updates/s; nominal 158 B/cell: %.2f of peak' % (
    st['stress_ms'] / n, alg['stress'] / (st['stress_ms'] / n) / 1e6, alg['stress'] / (st['stress_ms'] / n) / 1e6 / peak,
    st['particle_ms'] / n, alg['particle'] / (st['particle_ms'] / n) / 1e6, alg['particle'] / (st['particle_ms'] / n) / 1e6 / peak,
    st['run_ms'] / n, cells * n / st['run_ms'] / 1e6, 158.0 * cells * n / st['run_ms'] / 1e6 / peak))
