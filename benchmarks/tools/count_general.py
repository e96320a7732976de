import sys, os
sys.path.insert(0, ".")
from benchmarks.workloads import load_workload
from pywr_b200.engine import CudaEngine
for name, S in (("thames", 512), ("two_reservoir", 1024)):
    s, prog, info = load_workload(name, S, years=10)
    eng = CudaEngine(s, S, device=0)
    eng.load_program(prog)
    eng.run(0, prog.n_steps); eng.sync()
    st = eng.stats()
    n = S * prog.n_steps
    print(name, "pivots/step", st["pivots"] / n, "general solves/step", st["refactors"] / n)
    eng.close()
