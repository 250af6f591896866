"""Run a few time steps of one BASELINE config and print the per-kernel device times:
   python profiles/run_config.py <workload> [time_steps] [variant]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import FdtdSlab
name = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t0 = time.time()
w = workloads.make_workload(name, periods=int(os.environ.get('BB_PERIODS', '0')) or None)
print('workload', name, w['meta']['shape'], 'steps', w['meta']['steps'], 'built in %.1f s' % (time.time() - t0), flush=True)
kw = {k: v for k, v in w['kwargs'].items() if k not in ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')}
if os.environ.get('BB_PROF_ACC'):
    kw['SensorStart'] = 0
t0 = time.time()
s = FdtdSlab(*w['args'], kernel_variant=variant, **kw)
print('setup %.2f s, device bytes %.2f GB' % (time.time() - t0, s.stats()['device_bytes'] / 1e9), flush=True)
s.run(3)
st = s.run(n, profile=True)
cells = w['meta']['cells']
print('per step: stress %.3f ms  particle %.3f ms  other %.3f ms  total %.3f ms -> %.1f Gcell-updates/s' % (
    st['stress_ms'] / n, st['particle_ms'] / n, st['other_ms'] / n, st['run_ms'] / n, cells * n / st['run_ms'] / 1e6))
