"""GPU tests (-m gpu) at BASELINE.json's full sizes.  The oracle cannot run the whole batch in
seconds, so parity is checked (a) exactly on a subset of scenarios — scenarios are independent, so
column k of the full run must equal the oracle's run of scenario k alone — and (b) through
size-independent properties of the domain (volume bounds, supply never above demand, storage mass
balance, accumulators consistent with the time series they summarise)."""
import numpy as np
import pytest

from benchmarks.workloads import load_workload, synthetic_inflow, SEEDS, daily_index, WORKLOADS
from fixtures_util import rel_diff, run_program

pytestmark = pytest.mark.gpu


def _subset_oracle(name, S, picks, years):
    """Oracle run of the scenarios `picks` of the global batch (same per-scenario inflow generator)."""
    from oracle.oracle_engine import OracleEngine

    w = WORKLOADS[name]
    months = daily_index(w["start"], years=years).month.values
    cols = [synthetic_inflow(months, 1, SEEDS[w["seed"]], first_scenario=k)[:, 0] for k in picks]
    s, prog, _ = load_workload(name, len(picks), years=years, inflow=np.stack(cols, axis=1))
    eng = OracleEngine(s, len(picks), threads=8)
    outs, status, state = run_program(eng, prog)
    eng.close()
    return outs, status


@pytest.mark.parametrize("name,S,years", [("two_reservoir", 1024, 50), ("thames", 4096, 30)])
def test_full_size_subset_parity_and_properties(name, S, years):
    from pywr_b200.engine import CudaEngine

    s, prog, info = load_workload(name, S, years=years)
    T = prog.n_steps
    assert T == {50: 18262, 30: 10957}[years]                      # schedule-driven: bit-exact
    gpu = CudaEngine(s, S, device=0)
    outs, status, (x, vol, pc) = run_program(gpu, prog)
    assert status[0] == 0
    # (a) subset parity
    picks = [0, 1, S // 3, S // 2, S - 2, S - 1]
    ref, st_ref = _subset_oracle(name, S, picks, years)
    assert st_ref[0] == 0
    for r, (o, q) in enumerate(zip(outs, ref)):
        sub = o[:, picks] if o.ndim == 2 else o[picks]
        assert rel_diff(sub, q) <= 1e-9, (r, rel_diff(sub, q))
    # (b) properties
    kinds = list(prog.rec_kind)
    vol_recs = [r for r, k in enumerate(kinds) if k == 5]
    for r in vol_recs:
        v = outs[r]
        assert v.min() >= 0.0 and v.max() <= 200.0 + 1e-9           # min_volume 0, max_volume 200
        assert np.isfinite(v).all()
    flow_recs = [r for r, k in enumerate(kinds) if k == 1]
    for r in flow_recs:
        assert outs[r].min() >= -1e-9
    by_name = dict(zip(prog.rec_names, outs))
    if name == "two_reservoir":
        # supplied1 <= 1.8, supplied2 <= 1.0; DeficitFrequency counts the timesteps below demand
        sup1, sup2 = by_name["supplied1"], by_name["supplied2"]
        assert sup1.max() <= 1.8 + 1e-9 and sup2.max() <= 1.0 + 1e-9
        assert np.array_equal(by_name["deficit1"], (np.abs(sup1 - 1.8) > 1e-6).sum(axis=0).astype(float))
        assert np.array_equal(by_name["deficit2"], (np.abs(sup2 - 1.0) > 1e-6).sum(axis=0).astype(float))
        # final device state equals the last row of the volume recorders
        assert np.array_equal(vol[0], by_name["volume1"][-1]) and np.array_equal(vol[1], by_name["volume2"][-1])
    else:
        assert np.array_equal(vol[0], by_name["volume"][-1])
        assert by_name["supplied"].max() <= 1.7 * 1.2 + 1e-9      # demand_baseline * peak profile
    gpu.close()


def test_param_sets_batch_on_gpu():
    """Config 5 through the live Pywr objects on the GPU (skipped when Pywr is not importable)."""
    pytest.importorskip("pywr")
    from benchmarks.workloads import build_model_param_sets
    from oracle.oracle_engine import OracleEngine
    from pywr_b200.device_run import run_on_device

    P, I = 16, 8
    a = build_model_param_sets(P, I, years=2, solver="null")
    b = build_model_param_sets(P, I, years=2, solver="null")
    ra = run_on_device(a)
    rb = run_on_device(b, engine_factory=OracleEngine)
    for name in ra.recorders:
        assert rel_diff(ra.recorders[name], rb.recorders[name]) <= 1e-9, name
    assert ra.counters["pivots"] == rb.counters["pivots"]


def test_config5_full_batch_objectives_equal_oracle():
    """BASELINE.json configs[4] at its full batch: 256 control-curve parameter sets x 128 inflow scenarios = 32,768 scenario
    combinations (10 of the 30 years, so that the CPU oracle finishes in seconds and its arrays fit the host), through the live
    Pywr objects with arrays="none": the device keeps the temporal aggregates only and EVERY scenario's per-scenario recorder
    value equals the oracle's; the pivot counts are identical."""
    pytest.importorskip("pywr")
    from benchmarks.workloads import build_model_param_sets
    from oracle.oracle_engine import OracleEngine
    from pywr_b200.device_run import run_on_device

    P, I = 256, 128
    a = build_model_param_sets(P, I, years=10, solver="null")
    b = build_model_param_sets(P, I, years=10, solver="null")
    ra = run_on_device(a, arrays="none")
    rb = run_on_device(b, engine_factory=lambda s, S: OracleEngine(s, S, threads=16), arrays="none")
    assert ra.num_scenarios == P * I == 32768 and ra.timesteps == 3652
    for name in ra.recorders:
        assert ra.recorders[name].shape == (P * I,)
        assert rel_diff(ra.recorders[name], rb.recorders[name]) <= 1e-9, name
    assert ra.counters["pivots"] == rb.counters["pivots"]
    # objectives per parameter set the way the optimisation wrappers read them
    for rec_a, rec_b in zip(a.recorders, b.recorders):
        if hasattr(rec_a, "values"):
            assert np.allclose(np.asarray(rec_a.values()), np.asarray(rec_b.values()), rtol=1e-9, atol=1e-12)
    print(f"config 5: {ra.speed / 1e6:.0f} M scenario-timesteps/s on the device, read-out {ra.time_readout:.3f} s")
