"""Throughput of the device-resident run kernel versus scenario count (one GPU), for DESIGN.md /
profiles: python benchmarks/tools/scenario_scaling.py [workload] [years]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from benchmarks.workloads import load_workload  # noqa: E402
from pywr_b200.engine import CudaEngine  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "two_reservoir"
years = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = []
for S in (128, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
    s, prog, info = load_workload(workload, S, years=years)
    eng = CudaEngine(s, S, device=0)
    eng.load_program(prog)
    best = None
    for rep in range(4):
        eng.program_reset()
        eng.mark(0); eng.run(0, prog.n_steps); eng.mark(1)
        ms = eng.elapsed_ms()
        best = ms if best is None or ms < best else best
    st = eng.stats()
    rows.append({"scenarios": S, "timesteps": prog.n_steps, "kernel_ms": best, "variant": st["kernel_variant"],
                 "scenario_timesteps_per_s": S * prog.n_steps / best * 1e3,
                 "us_per_timestep": best * 1e3 / prog.n_steps, "failed": eng.program_status()[0]})
    print(json.dumps(rows[-1]), flush=True)
    eng.close()
