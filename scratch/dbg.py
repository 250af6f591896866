import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import FdtdSlab
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = workloads.make_workload('single_water', shape=(40, 44, 56), periods=5, pml=8)
kw = {k: v for k, v in w['kwargs'].items() if k not in ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')}
s = FdtdSlab(*w['args'], kernel_variant=variant, **kw)
print(s.run(n, profile=True))
