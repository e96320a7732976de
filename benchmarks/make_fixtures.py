"""Compiles the benchmark workloads into single-scenario device programs (needs baseline/_ref).
usage: python benchmarks/make_fixtures.py"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from benchmarks.workloads import FIXTURES, WORKLOADS, build_model  # noqa: E402
from pywr_b200.program import compile_program  # noqa: E402

if __name__ == "__main__":
    os.makedirs(FIXTURES, exist_ok=True)
    for name in WORKLOADS:
        model = build_model(name, 1, solver="null")
        model.setup()
        s, prog = compile_program(model, "route")
        s.save(os.path.join(FIXTURES, f"{name}.pstructure.npz"))
        prog.save(os.path.join(FIXTURES, f"{name}.program.npz"))
        print(name, f"{s.n_rows}x{s.n_cols} nnz={s.nnz} nodes={s.n_nodes} T={prog.n_steps} ops={prog.n_ops} "
              f"tables={prog.n_tables} recorders={prog.n_recorders} slots={s.n_slots}")
