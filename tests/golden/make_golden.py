"""Generates the golden fixtures under tests/golden/ (run in the BUILD container only).

Imports the reference framework built into baseline/_ref (baseline/build_ref.sh) and the reference's
model documents under /root/reference, runs each model with the CPU oracle plugged in as the Pywr
solver, and records, per timestep, the exact per-scenario input planes the solver gathered from the
model (slots), the timestep length and the oracle's outputs (node flows, column flows, objective),
together with the LP structure.  It also stores the optimal objective computed independently by
scipy's HiGHS for every recorded scenario-step.  The GPU box has neither /root/reference nor needs
pywr to replay these traces through the C-ABI.

usage: python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")

from pywr.model import Model  # noqa: E402

import oracle.pywr_oracle_solver as ops  # noqa: E402
from lp_numpy import highs_objective  # noqa: E402

REF = os.environ.get("PYWR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, path, formulation, max steps
    ("thames", "examples/thames/thames.json", "route", 90),
    ("thames", "examples/thames/thames.json", "edge", 40),
    ("two_reservoir", "examples/two_reservoir/two_reservoir.json", "route", 120),
    ("two_reservoir", "examples/two_reservoir/two_reservoir.json", "edge", 40),
    ("reservoir2", "tests/models/reservoir2.json", "route", 10),
    ("river_split_with_gauge1", "tests/models/river_split_with_gauge1.json", "route", 20),
    ("three_storage", "tests/models/three_storage.json", "route", 20),
    ("loss_link_parameter", "tests/models/loss_link_parameter.json", "route", 20),
    ("virtual_storage1", "tests/models/virtual_storage1.json", "route", 20),
    ("aggregated1", "tests/models/aggregated1.json", "route", 10),
    ("demand_saving2", "tests/models/demand_saving2.json", "route", 60),
    ("timeseries2", "tests/models/timeseries2.json", "route", 30),
    ("timeseries3", "tests/models/timeseries3.json", "route", 30),
    ("piecewise1", "tests/models/piecewise1.json", "edge", 5),
    # generated networks beyond the register-resident kernel variants (CTA-per-scenario kernel)
    ("network24", "generated:24", "edge", 25),
    ("network12", "generated:12", "route", 25),
]


class Recorder:
    def __init__(self):
        self.slots, self.days, self.node_flow, self.col_flow, self.objective = [], [], [], [], []


def run_case(name, path, formulation, max_steps):
    solver = "oracle" if formulation == "route" else "oracle-edge"
    if path.startswith("generated:"):
        from benchmarks.workloads import synthetic_network_document

        model = Model.load(synthetic_network_document(n_reaches=int(path.split(":")[1])), solver=solver)
    else:
        model = Model.load(os.path.join(REF, path), solver=solver)
    model.solver.save_routes_flows = True
    model.setup()
    rec = Recorder()
    sol = model.solver
    orig_step = sol._engine.step

    def step(days, slots, node_flow, col_flow, objective):
        n = orig_step(days, slots, node_flow, col_flow, objective)
        rec.slots.append(slots.copy()); rec.days.append(days)
        rec.node_flow.append(node_flow.copy()); rec.col_flow.append(col_flow.copy())
        rec.objective.append(objective[0].copy())
        return n

    sol._engine.step = step
    steps = min(max_steps, len(model.timestepper))
    for _ in range(steps):
        model.step()
    s = sol.structure
    slots = np.array(rec.slots)
    S = slots.shape[2]
    hobj = np.full((steps, S), np.nan)
    for t in range(steps):
        for k in range(S):
            v = highs_objective(s, slots[t], k, rec.days[t])
            hobj[t, k] = np.nan if v is None else v
    base = os.path.join(OUT, f"{name}.{formulation}")
    s.save(base + ".structure.npz")
    np.savez_compressed(base + ".trace.npz", slots=slots, days=np.array(rec.days), node_flow=np.array(rec.node_flow),
                        col_flow=np.array(rec.col_flow), objective=np.array(rec.objective), objective_highs=hobj)
    err = np.nanmax(np.abs(hobj - np.array(rec.objective)) / np.maximum(1.0, np.abs(hobj)))
    print(f"{name}.{formulation}: {s.n_rows}x{s.n_cols} nnz={s.nnz} slots={s.n_slots} S={S} T={steps} "
          f"max rel objective diff vs HiGHS = {err:.2e}")


if __name__ == "__main__":
    for case in CASES:
        run_case(*case)
