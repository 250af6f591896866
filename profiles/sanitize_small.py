"""Small end-to-end runs for compute-sanitizer: python profiles/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import FdtdSlab, collect_results
from BabelViscoFDTD.tools.RayleighAndBHTE import ForwardSimple
for name, shape, pml, over in (('ctx500_skull', (40, 44, 70), 6, {}), ('dome_stress', (36, 36, 40), 5, {}),
                               ('ctx500_skull', (30, 26, 36), 4, dict(SelMapsRMSPeakList=['ALLV', 'Vx', 'Sigmaxx', 'Sigmaxy', 'Pressure'], SelRMSorPeak=3))):
    w = workloads.make_workload(name, shape=shape, periods=2, pml=pml)
    kw = {k: v for k, v in w['kwargs'].items() if k not in ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')}
    kw.update(over)
    for variant in (0, 1):
        s = FdtdSlab(*w['args'], kernel_variant=variant, **kw)
        s.run()
        r = collect_results(s)
        s.close()
        print(name, shape, 'variant', variant, 'max RMS', float(max(v.max() for v in (r[1] or r[2]).values())))
rng = np.random.default_rng(0)
out = ForwardSimple(np.array(2000 + 0j).astype(np.complex64), rng.random((100, 3)).astype(np.float32), np.ones(100, np.float32),
                    np.ones(100, np.complex64), rng.random((777, 3)).astype(np.float32) + 2)
print('rayleigh', abs(out).max())
