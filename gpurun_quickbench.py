import sys, time, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from fixtures_util import *
from pywr_b200.engine import CudaEngine
s, prog, _, _ = load_program_case('two_reservoir.route')
for S in (1024, 8192, 32768):
    q = widen_program(s, prog, S, seed=1)
    eng = CudaEngine(s, S, device=0)
    eng.load_program(q)
    for rep in range(3):
        eng.program_reset()
        t0=time.perf_counter(); eng.run(0, q.n_steps); eng.sync(); dt=time.perf_counter()-t0
        st = eng.stats()
        print(f"S={S} T={q.n_steps} run {dt*1e3:.2f} ms  -> {S*q.n_steps/dt/1e6:.1f} M scenario-steps/s  pivots {st['pivots']} status {eng.program_status()}")
    eng.close()
