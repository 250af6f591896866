"""Where does the largest RMS pressure of the CTX-500 case sit, and is it steady?  Prints the sensor trace (last two periods)
at the voxel of the RMS maximum and at the focus-region maximum for the nominal run length and for twice that length.
   python profiles/run_edge_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import PropagationModel
for mult in (1, 2):
    w = workloads.make_workload('ctx500_skull', dense_sources=False, lean=True)
    m = w['meta']
    if mult > 1:
        w = workloads.make_workload('ctx500_skull', dense_sources=False, lean=True, periods=mult * m['steps'] // m['ppp'])
        m = w['meta']
    PM = PropagationModel()
    S, _, R, IP = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])
    rms = R['Pressure']
    MM = w['args'][0]
    idx = IP['IndexSensorMap'].astype(np.int64) - 1
    n1, n2, n3 = rms.shape
    def trace(v):
        f = v[0] + v[1] * n1 + v[2] * n1 * n2
        r = np.searchsorted(idx, f)
        return S['Pressure'][r]
    pk = np.unravel_index(int(np.argmax(rms)), rms.shape)
    inner = rms[40:-40, 40:-40, 100:-20]
    fo = np.unravel_index(int(np.argmax(inner)), inner.shape)
    fo = (fo[0] + 40, fo[1] + 40, fo[2] + 100)
    print('steps', m['steps'], 'RMS max %.4g at' % rms.max(), pk, 'label', int(MM[pk]), 'trace', np.array2string(trace(pk), precision=3))
    print('   interior max %.4g at' % rms[fo], fo, 'label', int(MM[fo]), 'trace', np.array2string(trace(fo), precision=3))
    # RMS maximum per material label outside / inside 4 cells of the PML
    edge = np.ones(rms.shape, bool); edge[16:-16, 16:-16, 16:-16] = False
    for lab in range(5):
        a = rms[(MM == lab) & edge]; b = rms[(MM == lab) & ~edge]
        print('   label %d: max RMS within 4 cells of the PML %.4g, elsewhere %.4g' % (lab, a.max() if a.size else 0, b.max() if b.size else 0))
