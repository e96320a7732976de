"""BASELINE.json configs[0] / SURVEY.md 8(d) config 1 as a committed fixture (BUILD container only; needs baseline/_ref and
/root/reference): the reference's examples/two_reservoir model, 1 scenario, DAILY timestep over its whole 100-year record.

  * config1/route.pstructure.npz / .program.npz — LP structure and the compiled device program of the run;
  * config1/route.expected.npz — from a HOST run of the reference framework (Model.run: the reference's own Parameter,
    Storage.after and Recorder code around the CPU oracle as LP solver): every 8th row of each array recorder plus its sum,
    min and max over the run, and the values of the accumulating recorders;
  * config1/route.structure.npz + route.trace.npz — the per-timestep (Solver.solve) LP structure and the solver's input planes at 96 timesteps spread over the run with the LP optimum of each from
    scipy's HiGHS (an independent simplex): the objective parity of north_star is pinned at these.

usage: python tests/golden/make_config1.py
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
from pywr.recorders import (DeficitFrequencyNodeRecorder, NumpyArrayNodeRecorder, NumpyArrayStorageRecorder,  # noqa: E402
                            TotalDeficitNodeRecorder, TotalFlowNodeRecorder)

import oracle.pywr_oracle_solver  # noqa: E402,F401
from lp_numpy import highs_objective  # noqa: E402
from pywr_b200.program import ARRAY_KINDS, compile_program  # noqa: E402

REF = os.environ.get("PYWR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
STRIDE, N_SAMPLES = 8, 96


def make():
    m = Model.load(os.path.join(REF, "examples/two_reservoir/two_reservoir.json"), solver="oracle")
    m.timestepper.delta = 1
    m.timestepper.end = "2199-12-30"          # the inflow record ends here
    recs = [NumpyArrayStorageRecorder(m, m.nodes["reservoir1"], name="vol_reservoir1"),
            NumpyArrayStorageRecorder(m, m.nodes["reservoir2"], name="vol_reservoir2"),
            NumpyArrayNodeRecorder(m, m.nodes["demand1"], name="flow_demand1"),
            NumpyArrayNodeRecorder(m, m.nodes["demand2"], name="flow_demand2"),
            NumpyArrayNodeRecorder(m, m.nodes["transfer"], name="flow_transfer"),
            TotalDeficitNodeRecorder(m, m.nodes["demand1"], name="tdef_demand1"),
            TotalFlowNodeRecorder(m, m.nodes["transfer"], name="tf_transfer"),
            DeficitFrequencyNodeRecorder(m, m.nodes["demand2"], name="df_demand2")]
    return m, recs


def main():
    m, recs = make()
    m.setup()
    T = len(m.timestepper)
    sample = set(np.unique(np.linspace(0, T - 1, N_SAMPLES).astype(int)).tolist())
    sol = m.solver
    orig_step = sol._engine.step
    trace = {"slots": [], "days": [], "objective": [], "t": []}
    counter = [0]

    def step(days, slots, node_flow, col_flow, objective):
        n = orig_step(days, slots, node_flow, col_flow, objective)
        if counter[0] in sample:
            trace["slots"].append(slots.copy()); trace["days"].append(days)
            trace["objective"].append(objective[0].copy()); trace["t"].append(counter[0])
        counter[0] += 1
        return n

    sol._engine.step = step
    m.run()
    assert counter[0] == T, (counter[0], T)
    expected = {r.name: (np.array(r.data) if hasattr(r, "data") else np.array(r.values())) for r in recs}
    s = sol.structure
    slots = np.array(trace["slots"])          # [n, n_slots, 1]
    hobj = np.array([highs_objective(s, slots[i], 0, trace["days"][i]) for i in range(len(trace["t"]))])
    m2, recs2 = make()
    m2.setup()
    s2, prog = compile_program(m2, "route", recorders=recs2)
    os.makedirs(os.path.join(OUT, "config1"), exist_ok=True)
    base = os.path.join(OUT, "config1", "route")
    s2.save(base + ".pstructure.npz")
    prog.save(base + ".program.npz")
    out = {"names": np.array([r.name for r in prog.recorders]), "stride": np.array(STRIDE)}
    for i, r in enumerate(prog.recorders):
        e = expected[r.name]
        if int(prog.rec_kind[i]) in ARRAY_KINDS:
            out[f"rec{i}"] = e[::STRIDE].copy()
            out[f"agg{i}"] = np.stack([np.cumsum(e, axis=0)[-1], e.min(axis=0), e.max(axis=0)])
        else:
            out[f"rec{i}"] = e
    np.savez_compressed(base + ".expected.npz", **out)
    s.save(base + ".structure.npz")           # the Solver.solve-path structure the recorded input planes belong to
    np.savez_compressed(base + ".trace.npz", slots=slots, days=np.array(trace["days"]), t=np.array(trace["t"]),
                        objective=np.array(trace["objective"]), objective_highs=hobj)
    err = np.nanmax(np.abs(hobj - np.array(trace["objective"])[:, 0]) / np.maximum(1.0, np.abs(hobj)))
    print(f"config1: T={T} S=1 LP {s.n_rows}x{s.n_cols}; {len(trace['t'])} sampled timesteps, max rel objective diff oracle vs HiGHS "
          f"= {err:.2e}; ops={prog.n_ops} tables={prog.n_tables} recorders={prog.n_recorders}")


if __name__ == "__main__":
    main()
