"""Where does a half-step launch spend its time?  Per-CTA start/end times (globaltimer) of the last particle launch,
aggregated by tile class.   BB_CTA_TIMING=1 python profiles/cta_times.py [workload] [steps]"""
import os, sys
os.environ.setdefault('BB_CTA_TIMING', 'particle')   # or 'stress'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from babelbrain_b200 import workloads, _capi
from babelbrain_b200.propagation import FdtdSlab
name = sys.argv[1] if len(sys.argv) > 1 else 'ctx500_skull'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w = workloads.make_workload(name, periods=None if name == 'ctx500_skull' else 2)
kw = {k: v for k, v in w['kwargs'].items() if k not in ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')}
kw['SensorStart'] = 10 ** 6 // kw['SensorSubSampling']
s = FdtdSlab(*w['args'], **kw)
s.run(n)
n1, n2, n3 = w['meta']['shape']
ntk, ntj = (n3 + 63) // 64, (n2 + 7) // 8
rec = np.zeros((65536, 4), np.uint64)
_capi.check(s._L.bb_fdtd_debug_cta_times(s._h, _capi.ptr(rec), 65536))
rec = rec[rec[:, 1] > 0]
t0 = rec[:, 0].min()
start, end = (rec[:, 0] - t0).astype(np.float64) / 1e3, (rec[:, 1] - t0).astype(np.float64) / 1e3
bx, by, bz = (rec[:, 2] & 0xFFFFF).astype(int), ((rec[:, 2] >> 20) & 0xFFFFF).astype(int), (rec[:, 2] >> 40).astype(int)
planes = rec[:, 3].astype(np.float64)
per_plane = (end - start) / planes
P = w['meta']['pml']
kedge = (bx * 64 < P) | (bx * 64 + 64 > n3 - P)
jedge = (by * 8 < P) | (by * 8 + 8 > n2 - P)
print('%s: %d CTAs of the last %s launch, makespan %.1f us' % (name, len(rec), os.environ['BB_CTA_TIMING'], end.max()))
for label, m in (('interior tiles', ~kedge & ~jedge), ('k-PML tiles', kedge & ~jedge), ('j-PML tiles', jedge & ~kedge), ('corner tiles', kedge & jedge)):
    if m.any():
        print('  %-15s n=%5d  us/plane: mean %.3f  p10 %.3f  p90 %.3f   (chunks >= 32 planes: %.3f)' % (
            label, m.sum(), per_plane[m].mean(), np.percentile(per_plane[m], 10), np.percentile(per_plane[m], 90),
            per_plane[m & (planes >= 32)].mean() if (m & (planes >= 32)).any() else float('nan')))
for z in sorted(set(bz)):
    m = bz == z
    print('  chunk %2d: planes %3d  CTAs %4d  start %.1f..%.1f us  end %.1f..%.1f us  us/plane %.3f' % (z, planes[m][0], m.sum(), start[m].min(), start[m].max(), end[m].min(), end[m].max(), per_plane[m].mean()))
busy = (end - start).sum()
print('  sum of CTA times / (148 x makespan) = %.3f' % (busy / (148 * end.max())))
