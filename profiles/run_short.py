"""Short run of the bench workload for ncu captures: python profiles/run_short.py [time_steps] [variant]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import FdtdSlab
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
name = sys.argv[3] if len(sys.argv) > 3 else 'ctx500_skull'
w = workloads.make_workload(name)
kw = {k: v for k, v in w['kwargs'].items() if k not in ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')}
if os.environ.get('BB_PROF_ACC'):
    kw['SensorStart'] = 0   # put the RMS window in range so the RMS-accumulating kernel variants are the ones captured
s = FdtdSlab(*w['args'], kernel_variant=variant, **kw)
print(s.run(n, profile=True))
