"""Generates the large-network fixture of SURVEY.md 8(d) config 4 (run in the BUILD container only:
it imports the reference framework from baseline/_ref and scipy's HiGHS).

A generated river network of ~2,000 nodes (benchmarks.workloads.synthetic_network_document with 600
reaches: catchments, river reaches, tributary reservoirs, demands, MRF gauges, a transfer and an
aggregated licence) is run in the edge formulation with a HiGHS-backed stand-in solver; recorded per
timestep: the input planes the solver gathers (slots), the timestep length and HiGHS's optimal
objective.  Output: benchmarks/fixtures/network600.edge.{structure,trace}.npz

usage: python tests/golden/make_network_fixture.py [n_reaches] [steps]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)

import pywr.solvers as ps  # noqa: E402
from pywr.model import Model  # noqa: E402

from benchmarks.workloads import synthetic_network_document  # noqa: E402
from oracle.pdhg_oracle import row_bounds  # noqa: E402
from pywr_b200.solver import BatchedLPSolver  # noqa: E402


class HighsEngine:
    """Duck-typed engine (same interface as CudaEngine) that solves scenario 0 with HiGHS."""

    def __init__(self, s, S):
        self.s, self.S = s, S
        self.A = sp.csr_matrix((s.a_val, s.a_col, s.a_indptr), shape=(s.n_rows, s.n_cols))
        self.F = sp.csr_matrix((s.flow_w, s.flow_col, s.flow_indptr), shape=(s.n_nodes, s.n_cols))
        self.slots, self.days, self.objective = [], [], []

    def alloc_planes(self, n):
        return np.zeros((n, self.S))

    def reset(self, v=None):
        pass

    def set_consts(self, c):
        pass

    def close(self):
        pass

    def status(self):
        return -1, 0

    def step(self, days, slots, node_flow, col_flow, objective):
        s = self.s
        lo = np.empty(s.n_rows); hi = np.empty(s.n_rows)
        for i in range(s.n_rows):
            lo[i], hi[i] = row_bounds(s, i, slots[:, 0], days, None)
        c = np.zeros(s.n_cols)
        for j in range(s.n_cols):
            for k in range(s.cost_indptr[j], s.cost_indptr[j + 1]):
                r = s.cost_ref[k]
                c[j] += s.cost_w[k] * (slots[r, 0] if r >= 0 else s.consts[-r - 1])
        eq = lo == hi
        ub, lb = np.isfinite(hi) & ~eq, np.isfinite(lo) & ~eq
        res = linprog(c, A_ub=sp.vstack([self.A[ub], -self.A[lb]]), b_ub=np.concatenate([hi[ub], -lo[lb]]),
                      A_eq=self.A[eq], b_eq=lo[eq], bounds=(0, None), method="highs")
        assert res.status == 0, res.message
        self.slots.append(slots.copy()); self.days.append(days); self.objective.append(res.fun)
        node_flow[:, 0] = self.F @ res.x
        return 0


class HighsSolver(BatchedLPSolver):
    name = "highs-capture"
    formulation = "edge"

    def _make_engine(self, s, S):
        return HighsEngine(s, S)


def main():
    n_reaches = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    ps.solver_registry.append(HighsSolver)
    model = Model.load(synthetic_network_document(n_reaches=n_reaches, end="2100-12-31"), solver="highs-capture")
    model.setup()
    for _ in range(steps):
        model.step()
    eng, s = model.solver._engine, model.solver._structure
    out = os.path.join(ROOT, "benchmarks", "fixtures", f"network{n_reaches}.edge")
    s.save(out + ".structure.npz")
    np.savez_compressed(out + ".trace.npz", slots=np.array(eng.slots), days=np.array(eng.days),
                        objective_highs=np.array(eng.objective)[:, None])
    print(f"{len(model.nodes)} nodes, LP {s.n_rows} x {s.n_cols}, nnz {s.nnz}, {s.n_slots} slots, {steps} steps -> {out}.*")


if __name__ == "__main__":
    main()
