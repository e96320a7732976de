"""Per-phase cycle counts of the device-resident run kernel (scenario 0), from a library built with
-DB200LP_PHASE_TIMING:
  nvcc ... -DB200LP_PHASE_TIMING -o pywr_b200/libb200lp_timing.so pywr_b200/csrc/b200lp.cu
  gpurun -- python benchmarks/tools/phase_timing.py
"""
import sys, ctypes, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pywr_b200.engine as E
E.LIB_PATH = 'pywr_b200/libb200lp_timing.so'
from benchmarks.workloads import load_workload
s,q,info = load_workload('two_reservoir', 1024, years=2)
L = E.load_library(E.LIB_PATH)
eng = E.CudaEngine(s, 1024, device=0)
eng.load_program(q)
eng.run(0, q.n_steps); eng.sync()
out = (ctypes.c_ulonglong*16)()
L.b200lp_debug_phase_cycles(out)
v = np.array(list(out), dtype=np.float64) / q.n_steps
names = ['tables+sync','ops','assemble','solve','extract','storage','recorders']
print('cycles per step per phase (scenario 0):')
for n,x in zip(names, v): print(f'  {n:12s} {x:8.1f}')
print('  total', v.sum())
