"""Generates the device-run fixture of SURVEY.md 8(d) config 4 (BUILD container only: imports the reference framework
from baseline/_ref and, as test infrastructure, the oracle).

The generated 2,073-node river network (benchmarks.workloads.synthetic_network_document, 600 reaches) gets a scenario axis
`inflow`: every catchment inflow is its monthly profile times a per-scenario wetness factor (ConstantScenarioParameter), so
the whole parameter set is device-evaluable.  Written: (1) the compiled program + LP structure (edge formulation) for the
C-ABI, (2) the recorder arrays of a HOST run of the reference framework (Model.run, the reference's own Parameter /
Storage.after / Recorder code) with the host build of the revised-simplex core as solver.

Output: benchmarks/fixtures/network600.{pstructure,program,expected}.npz
usage: python tests/golden/make_network_program.py [n_reaches] [end-date]
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from pywr.model import Model  # noqa: E402
from pywr.recorders import (DeficitFrequencyNodeRecorder, MeanFlowNodeRecorder, MinimumVolumeStorageRecorder,  # noqa: E402
                            NumpyArrayNodeRecorder, NumpyArrayStorageRecorder, TotalDeficitNodeRecorder)

import oracle.pywr_oracle_solver  # noqa: E402,F401
from benchmarks.workloads import synthetic_network_document  # noqa: E402
from pywr_b200.program import compile_program  # noqa: E402

WETNESS = [1.0, 0.6]


def document(n_reaches, end):
    doc = synthetic_network_document(n_reaches=n_reaches, end=end)
    doc["scenarios"] = [{"name": "inflow", "size": len(WETNESS)}]
    params = doc["parameters"]
    params["wet"] = {"type": "constantscenario", "scenario": "inflow", "values": WETNESS}
    for name in [k for k in params if k.startswith("inflow")]:
        params[name] = {"type": "aggregated", "agg_func": "product", "parameters": [params[name], "wet"]}
    return doc


def add_recorders(m):
    recs = []
    stores = [n for n in m.nodes if type(n).__name__ == "Storage"]
    towns = [n for n in m.nodes if n.name.startswith("town")]
    pick = lambda xs, k: [xs[i] for i in np.linspace(0, len(xs) - 1, k).astype(int)]  # noqa: E731
    for n in pick(stores, 6):
        recs += [NumpyArrayStorageRecorder(m, n, name=f"vol_{n.name}"), MinimumVolumeStorageRecorder(m, n, name=f"minv_{n.name}")]
    for n in pick(towns, 8):
        recs += [TotalDeficitNodeRecorder(m, n, name=f"tdef_{n.name}"), DeficitFrequencyNodeRecorder(m, n, name=f"df_{n.name}")]
    recs.append(NumpyArrayNodeRecorder(m, m.nodes["sea"], name="flow_sea"))
    recs.append(MeanFlowNodeRecorder(m, m.nodes["transfer"], name="mf_transfer"))
    return recs


def main():
    n_reaches = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    end = sys.argv[2] if len(sys.argv) > 2 else "2100-12-31"
    m = Model.load(document(n_reaches, end), solver="oracle-rsx-edge")
    recs = add_recorders(m)
    t0 = time.time()
    m.run()
    print(f"host run: {time.time() - t0:.1f} s, {m.solver._engine.stats()}")
    expected = {r.name: (np.array(r.data) if hasattr(r, "data") else np.array(r.values())) for r in recs}
    m2 = Model.load(document(n_reaches, end), solver="oracle-rsx-edge")
    recs2 = add_recorders(m2)
    m2.setup()
    s, prog = compile_program(m2, "edge", recorders=recs2)
    base = os.path.join(ROOT, "benchmarks", "fixtures", f"network{n_reaches}")
    s.save(base + ".pstructure.npz")
    prog.save(base + ".program.npz")
    np.savez_compressed(base + ".expected.npz", names=np.array([r.name for r in prog.recorders]),
                        **{f"rec{i}": expected[r.name] for i, r in enumerate(prog.recorders)})
    print(f"network{n_reaches}: {len(m2.nodes)} nodes, LP {s.n_rows} x {s.n_cols}, T={prog.n_steps} S={len(m2.scenarios.combinations)} "
          f"ops={prog.n_ops} tables={prog.n_tables} recorders={prog.n_recorders} slots={s.n_slots}")


if __name__ == "__main__":
    main()
