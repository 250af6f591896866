"""End-to-end phases of the public call, repeated: python profiles/run_e2e.py [repeats]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import PropagationModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
w = workloads.make_workload('ctx500_skull')
for it in range(n):
    PM = PropagationModel()
    t0 = time.perf_counter()
    res = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])
    t1 = time.perf_counter()
    del res
    print('call %d: %.3f s  phases %s  (del %.3f s)' % (it, t1 - t0, {k: round(v, 3) for k, v in PM.last_timing.items() if k.endswith('_s')}, time.perf_counter() - t1), flush=True)
