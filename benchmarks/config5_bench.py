"""the optimisation-style batch.  two_reservoir, 256 control-curve parameter sets
(Scenario "params", ScenarioMonthlyProfileParameter) x 128 inflow scenarios flattened into one batch of
32,768 scenario combinations, 30 years daily, through the reference framework's own objects:
``pywr_b200.device_run.run_on_device(model)`` (needs the reference framework: baseline/_ref).

Prints one JSON line: scenario-timesteps/s of the device run, compile / read-out times, and the objective
of every parameter set in the reference's ``evaluate`` sense (pywr/optimisation/platypus.py:115-154: the
recorder's value aggregated over the inflow scenarios of that parameter set).

usage: python benchmarks/config5_bench.py [--param-sets 256] [--inflow 128] [--years 30]
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
    ap.add_argument("--param-sets", type=int, default=256)
    ap.add_argument("--inflow", type=int, default=128)
    ap.add_argument("--years", type=int, default=30)
    ap.add_argument("--arrays", default="none", choices=["eager", "lazy", "none"],
                    help="none (default): objectives from the temporal aggregates the device keeps, no [T, S] array is stored or "
                         "shipped; eager: every array into Pywr's recorder objects (round 1: 11 GB, 4.1 s)")
    a = ap.parse_args()
    from benchmarks.workloads import build_model_param_sets
    from pywr_b200.device_run import run_on_device

    t0 = time.perf_counter()
    model = build_model_param_sets(a.param_sets, a.inflow, years=a.years, solver="null")
    t1 = time.perf_counter()
    res = run_on_device(model, arrays=a.arrays)
    t2 = time.perf_counter()
    P, I = a.param_sets, a.inflow
    combos = np.array([np.asarray(si.indices) for si in model.scenarios.combinations])
    names = [sc.name for sc in model.scenarios.scenarios]
    pa = names.index("params")
    objectives = {}
    for name, values in res.recorders.items():
        v = np.asarray(values)                                      # arrays="none": already the temporal aggregate per scenario
        per_scenario = v.mean(axis=0) if v.ndim == 2 else v          # temporal mean of the array recorders
        per_set = np.array([per_scenario[combos[:, pa] == p].mean() for p in range(P)])
        objectives[name] = {"min": float(per_set.min()), "max": float(per_set.max()), "argbest": int(per_set.argmin())}
    print(json.dumps({
        "metric": "scenario-timesteps/sec (config 5: 256 parameter sets x 128 inflow scenarios)", "value": res.speed,
        "unit": "scenario-timesteps/s", "scenarios": res.num_scenarios, "timesteps": res.timesteps,
        "param_sets": P, "inflow_scenarios": I, "run_s": res.time_run, "compile_s": res.time_compile,
        "readout_s": res.time_readout, "model_build_s": t1 - t0, "total_s": t2 - t0, "arrays": a.arrays,
        "end_to_end_value": res.num_scenarios * res.timesteps / (res.time_compile + res.time_run + res.time_readout),
        "end_to_end_note": "compile + device run + read-out of every recorder's per-scenario value (objectives), Pywr setup excluded",
        "pivots_per_scenario_timestep": res.counters["pivots"] / (res.num_scenarios * res.timesteps),
        "objectives_per_parameter_set": objectives}))


if __name__ == "__main__":
    main()
