"""Two-GPU run of a small grid through the public call (slabs as threads of one process, halos pushed over NVLink) for
compute-sanitizer:  compute-sanitizer --tool memcheck python profiles/sanitize_two_gpus.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from babelbrain_b200 import workloads
from BabelViscoFDTD.PropagationModel import PropagationModel
w = workloads.make_workload('ctx500_skull', shape=(40, 44, 70), periods=2, pml=6)
one = PropagationModel().StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])
two = PropagationModel().StaggeredFDTD_3D_with_relaxation(*w['args'], NumberGPUs=2, **w['kwargs'])
a, b = one[2]['Pressure'], two[2]['Pressure']
print('two slabs vs one GPU: RMS rel-L2 %.3g, sensors rel-L2 %.3g' % (
    np.linalg.norm(a - b) / np.linalg.norm(a), np.linalg.norm(one[0]['Pressure'] - two[0]['Pressure']) / np.linalg.norm(one[0]['Pressure'])))
