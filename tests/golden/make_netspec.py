"""Fixtures of SURVEY.md 8(d) config 4 AS SPECIFIED (BUILD container only: imports the reference framework from
baseline/_ref, scipy's HiGHS and, as test infrastructure, the oracle).

The generated network (benchmarks.workloads.spec_network_document: ~2,000 primitive nodes at scale 1 — 150 reservoirs, 250
catchments, 300 demands, 40 AggregatedNodes of which 20 with factors, 60 PiecewiseLinks, a river tree plus transfers that close
cycles; edge formulation) with per-scenario DAILY stochastic inflows (add_stochastic_inflows).  Written for `name`:
  <name>.pstructure.npz / .program.npz  the compiled device program;
  <name>.expected.npz                   recorder arrays of a HOST run of the reference framework (Model.run: the reference's own
                                        Parameter / Storage.after / Recorder code) whose LP solver is the DICTIONARY oracle
                                        (oracle/lp_oracle.c) - independent of the revised simplex the CUDA path uses;
  <name>.edge.structure.npz / .edge.trace.npz  the Solver.solve-path structure, the input planes of the first `trace` timesteps
                                        and HiGHS's optimum of every recorded scenario-timestep.

usage: python tests/golden/make_netspec.py small | full
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")

from pywr.model import Model  # noqa: E402
from pywr.recorders import (DeficitFrequencyNodeRecorder, MeanFlowNodeRecorder, MinimumVolumeStorageRecorder,  # noqa: E402
                            NumpyArrayNodeRecorder, NumpyArrayStorageRecorder, TotalDeficitNodeRecorder)

import oracle.pywr_oracle_solver  # noqa: E402,F401
from benchmarks.workloads import add_stochastic_inflows, spec_network_document  # noqa: E402
from lp_numpy import highs_objective  # noqa: E402
from pywr_b200.program import compile_program  # noqa: E402

VARIANTS = {
    # name: (scale, scenarios, end date, traced timesteps, output base)
    "small": (0.12, 3, "2100-04-29", 20, os.path.join(ROOT, "tests", "golden", "netspec", "small")),
    "full": (1.0, 2, "2100-12-31", 6, os.path.join(ROOT, "benchmarks", "fixtures", "netspec2000")),
}


SOLVER = os.environ.get("NETSPEC_SOLVER", "oracle-edge")   # oracle-rsx-edge: host build of the revised simplex (minutes, not an hour)


def build(scale, S, end):
    m = Model.load(spec_network_document(scale=scale, end=end), solver=SOLVER)
    add_stochastic_inflows(m, S)
    return m


def add_recorders(m):
    recs = []
    pick = lambda xs, k: [xs[i] for i in sorted(set(np.linspace(0, len(xs) - 1, k).astype(int)))]  # noqa: E731
    stores = sorted((n for n in m.nodes if type(n).__name__ == "Storage"), key=lambda n: n.name)
    towns = sorted((n for n in m.nodes if n.name.startswith("town")), key=lambda n: n.name)
    farms = sorted((n for n in m.nodes if n.name.startswith("farm")), key=lambda n: n.name)
    pws = sorted((n for n in m.nodes if type(n).__name__ == "PiecewiseLink"), key=lambda n: n.name)
    for n in pick(stores, 6):
        recs += [NumpyArrayStorageRecorder(m, n, name=f"vol_{n.name}"), MinimumVolumeStorageRecorder(m, n, name=f"minv_{n.name}")]
    for n in pick(towns, 6):
        recs += [TotalDeficitNodeRecorder(m, n, name=f"tdef_{n.name}"), DeficitFrequencyNodeRecorder(m, n, name=f"df_{n.name}")]
    for n in pick(farms, 4):      # farms sit behind the ratio-constrained intakes
        recs.append(NumpyArrayNodeRecorder(m, n, name=f"flow_{n.name}"))
    for n in pick(pws, 2):
        recs.append(NumpyArrayNodeRecorder(m, n, name=f"flow_{n.name}"))
    recs.append(NumpyArrayNodeRecorder(m, m.nodes["sea"], name="flow_sea"))
    recs.append(MeanFlowNodeRecorder(m, m.nodes["transfer0"], name="mf_transfer0"))
    return recs


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    scale, S, end, n_trace, base = VARIANTS[which]
    if SOLVER != "oracle-edge":
        base += "_" + SOLVER.replace("oracle-", "").replace("-edge", "")
    os.makedirs(os.path.dirname(base), exist_ok=True)
    m = build(scale, S, end)
    recs = add_recorders(m)
    m.setup()
    sol = m.solver
    trace = {"slots": [], "days": [], "objective": []}
    orig_step = sol._engine.step

    def step(days, slots, node_flow, col_flow, objective):
        n = orig_step(days, slots, node_flow, col_flow, objective)
        if len(trace["days"]) < n_trace:
            trace["slots"].append(slots.copy()); trace["days"].append(days); trace["objective"].append(objective[0].copy())
        return n

    sol._engine.step = step
    t0 = time.time()
    m.run()
    eng = sol._engine
    print(f"host run ({SOLVER}): {time.time() - t0:.1f} s, {eng.counters() if hasattr(eng, 'counters') else eng.stats()}")
    expected = {r.name: (np.array(r.data) if hasattr(r, "data") else np.array(r.values())) for r in recs}
    s = sol.structure
    slots = np.array(trace["slots"])
    t0 = time.time()
    hobj = np.array([[highs_objective(s, slots[t], k, trace["days"][t]) for k in range(S)] for t in range(len(trace["days"]))])
    err = np.nanmax(np.abs(hobj - np.array(trace["objective"])) / np.maximum(1.0, np.abs(hobj)))
    print(f"HiGHS on {hobj.size} scenario-timesteps: {time.time() - t0:.1f} s, max rel objective diff oracle vs HiGHS {err:.2e}")
    s.save(base + ".edge.structure.npz")
    np.savez_compressed(base + ".edge.trace.npz", slots=slots, days=np.array(trace["days"]), objective=np.array(trace["objective"]),
                        objective_highs=hobj)
    m2 = build(scale, S, end)
    recs2 = add_recorders(m2)
    m2.setup()
    s2, prog = compile_program(m2, "edge", recorders=recs2)
    s2.save(base + ".pstructure.npz")
    prog.save(base + ".program.npz")
    np.savez_compressed(base + ".expected.npz", names=np.array([r.name for r in prog.recorders]),
                        **{f"rec{i}": expected[r.name] for i, r in enumerate(prog.recorders)})
    kinds = {}
    for n in m2.nodes:
        kinds[type(n).__name__] = kinds.get(type(n).__name__, 0) + 1
    print(f"{which}: {len(list(m2.graph.nodes()))} graph nodes {kinds}, LP {s2.n_rows} x {s2.n_cols} nnz {s2.nnz}, T={prog.n_steps} S={S} "
          f"ops={prog.n_ops} tables={prog.n_tables} recorders={prog.n_recorders} slots={s2.n_slots}")


if __name__ == "__main__":
    main()
