"""SURVEY.md 8(d) config 4 as a Pywr USER runs it: the generated 2,073-node network is built as a ``pywr.Model`` (reference
framework, baseline/_ref), given an `inflow` scenario axis of S wetness factors, and run with
``pywr_b200.device_run.run_on_device(model, formulation="edge")``: Pywr's own setup / reset, the model compiled into a device
program, the whole year on the GPU (revised simplex of the large-LP path), recorder arrays written back into Pywr's recorder
objects.  Prints one JSON line with every phase timed.

usage: python benchmarks/config4_pywr_bench.py [--scenarios 8192] [--reaches 600] [--end 2100-12-31]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(_REF, "pywr")):
    sys.path.insert(0, _REF)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=8192)
    ap.add_argument("--reaches", type=int, default=600)
    ap.add_argument("--end", default="2100-12-31")
    a = ap.parse_args()
    from pywr.model import Model
    from pywr.recorders import (DeficitFrequencyNodeRecorder, MinimumVolumeStorageRecorder, NumpyArrayNodeRecorder,
                                NumpyArrayStorageRecorder, TotalDeficitNodeRecorder)

    from benchmarks.workloads import synthetic_network_document
    from pywr_b200.device_run import run_on_device

    S = a.scenarios
    rng = np.random.default_rng(20240903)
    wet = rng.uniform(0.5, 1.25, size=S)
    wet[:2] = [1.0, 0.6]
    t0 = time.perf_counter()
    doc = synthetic_network_document(n_reaches=a.reaches, end=a.end)
    doc["scenarios"] = [{"name": "inflow", "size": S}]
    params = doc["parameters"]
    params["wet"] = {"type": "constantscenario", "scenario": "inflow", "values": [float(v) for v in wet]}
    for name in [k for k in params if k.startswith("inflow")]:
        params[name] = {"type": "aggregated", "agg_func": "product", "parameters": [params[name], "wet"]}
    m = Model.load(doc, solver="null")
    stores = [n for n in m.nodes if type(n).__name__ == "Storage"]
    towns = [n for n in m.nodes if n.name.startswith("town")]
    pick = lambda xs, k: [xs[i] for i in np.linspace(0, len(xs) - 1, k).astype(int)]  # noqa: E731
    for n in pick(stores, 6):
        NumpyArrayStorageRecorder(m, n, name=f"vol_{n.name}")
        MinimumVolumeStorageRecorder(m, n, name=f"minv_{n.name}")
    for n in pick(towns, 8):
        TotalDeficitNodeRecorder(m, n, name=f"tdef_{n.name}")
        DeficitFrequencyNodeRecorder(m, n, name=f"df_{n.name}")
    NumpyArrayNodeRecorder(m, m.nodes["sea"], name="flow_sea")
    t1 = time.perf_counter()
    res = run_on_device(m, formulation="edge")
    t2 = time.perf_counter()
    sea = np.asarray(m.recorders["flow_sea"].data)
    tdef = {r.name: np.asarray(r.values()) for r in m.recorders if r.name.startswith("tdef_")}
    print(json.dumps({
        "metric": "scenario-timesteps/sec (config 4 through run_on_device(model))", "value": res.speed, "unit": "scenario-timesteps/s",
        "scenarios": res.num_scenarios, "timesteps": res.timesteps, "nodes": len(m.nodes),
        "model_build_s": t1 - t0, "setup_s": res.time_setup, "compile_s": res.time_compile, "run_s": res.time_run,
        "readout_s": res.time_readout, "total_s": t2 - t0, "end_to_end_value": res.num_scenarios * res.timesteps / (t2 - t1),
        "pivots_per_scenario_timestep": res.counters["pivots"] / (res.num_scenarios * res.timesteps),
        "flow_sea_mean_wettest_driest": [float(sea[:, int(np.argmax(wet))].mean()), float(sea[:, int(np.argmin(wet))].mean())],
        "total_deficit_mean_over_scenarios": {k: float(v.mean()) for k, v in tdef.items()}}))


if __name__ == "__main__":
    main()
