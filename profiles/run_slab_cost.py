"""What do the planes of the absorbing layer along axis 0 cost?  One GPU runs, alone and without halo exchange
(BB_EXPERIMENT_NOHALO=1), slabs of a decomposed grid: the first slab (holds the 12 layer planes) and an interior slab with
the same number of planes:   BB_EXPERIMENT_NOHALO=1 python profiles/run_slab_cost.py <workload> <n1> <n2> <n3> <nslabs> [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import FdtdSlab
from babelbrain_b200.slab import SlabPlan
name, n1, n2, n3, nsl = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
steps = int(sys.argv[6]) if len(sys.argv) > 6 else 100
plan = SlabPlan(n1, nsl)
for rank in sorted({0, nsl // 2, nsl - 1}):
    glo, ghi = plan.with_halo(rank)
    w = workloads.make_workload(name, shape=(n1, n2, n3), periods=4, planes=(glo, ghi), lean=True, dense_sources=False)
    kw = {k: v for k, v in w['kwargs'].items() if k not in ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')}
    s = FdtdSlab(*w['args'], rank=rank, nranks=nsl, origin=glo, n1_global=n1, **kw)
    n = min(steps, w['meta']['steps'] - 5)
    s.run(5)
    st = s.run(n, profile=True)
    print('slab %d planes [%d,%d): stress %.4f ms  particle %.4f ms  other %.4f ms  step %.4f ms' % (
        rank, s.i0, s.i1, st['stress_ms'] / n, st['particle_ms'] / n, st['other_ms'] / n, st['run_ms'] / n), flush=True)
    s.close() if hasattr(s, 'close') else None
    del s
