"""The benchmark workloads: the numpy-only program construction used by bench.py must be the program
the compiler produces from the live Pywr model, and the C-ABI run must match the reference host run."""
import numpy as np
import pytest

from benchmarks.workloads import WORKLOADS, algorithmic_bytes_per_scenario_step, load_workload, synthetic_inflow
from fixtures_util import rel_diff, run_program
from oracle.oracle_engine import OracleEngine


def test_synthetic_inflow_is_rank_invariant():
    months = np.tile(np.arange(1, 13), 10)
    whole = synthetic_inflow(months, 8, 123)
    lo = synthetic_inflow(months, 4, 123, first_scenario=0)
    hi = synthetic_inflow(months, 4, 123, first_scenario=4)
    assert np.array_equal(whole[:, :4], lo) and np.array_equal(whole[:, 4:], hi)
    assert np.all(whole > 0) and np.isfinite(whole).all()


def test_algorithmic_bytes_match_survey():
    _, _, info = load_workload("two_reservoir", 4, years=1)
    # 906 B with 4 recorders; this workload records 7 values per scenario-timestep
    assert algorithmic_bytes_per_scenario_step(info) == 906 + 8 * (info["n_recorders"] - 4)
    _, _, info = load_workload("thames", 4, years=1)
    assert algorithmic_bytes_per_scenario_step(info) == 474 + 8 * (info["n_recorders"] - 2)


@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_fixture_program_equals_compiled_model(name):
    pytest.importorskip("pywr")
    import oracle.pywr_oracle_solver  # noqa: F401
    from benchmarks.workloads import build_model
    from pywr_b200.program import compile_program

    S, years = 6, 1
    model = build_model(name, S, years=years, solver="oracle")
    model.setup()
    s_ref, p_ref = compile_program(model, "route")
    s, p, info = load_workload(name, S, years=years)
    assert (s.n_rows, s.n_cols, s.n_slots) == (s_ref.n_rows, s_ref.n_cols, s_ref.n_slots)
    assert np.array_equal(s.dense(), s_ref.dense())
    assert p.n_steps == p_ref.n_steps and np.array_equal(p.op_kind, p_ref.op_kind)
    assert np.array_equal(p.op_args, p_ref.op_args) and np.array_equal(p.rec_kind, p_ref.rec_kind)
    assert np.array_equal(p.table_rows, p_ref.table_rows) and np.array_equal(p.table_cols, p_ref.table_cols)
    assert np.array_equal(p.table_data, p_ref.table_data)
    assert np.array_equal(p.initial_volume, p_ref.initial_volume)
    # and the run itself against the reference framework's host run
    a = OracleEngine(s, S); b = OracleEngine(s_ref, S)
    oa, sta, _ = run_program(a, p); ob, stb, _ = run_program(b, p_ref)
    assert sta == stb and sta[0] == 0
    for x, y in zip(oa, ob):
        assert np.array_equal(x, y)
    model2 = build_model(name, S, years=years, solver="oracle")
    model2.run()
    T = p.n_steps
    for r, rec in enumerate(model2.recorders):
        if type(rec).__name__ == "AggregatedRecorder":
            continue
        i = [q.name for q in p_ref.recorders].index(rec.name)
        host = np.array(rec.data) if hasattr(rec, "data") else np.array(rec.values())
        dev = oa[i] / T if type(rec).__name__ in ("DeficitFrequencyNodeRecorder", "MeanFlowNodeRecorder") else oa[i]
        assert rel_diff(dev, host) <= 1e-12, rec.name
    a.close(); b.close()


def test_param_sets_batch_matches_host_run():
    """Config 5: parameter sets flattened into a scenario axis; objectives per parameter set are read
    back in the reference's combination order (params-major, inflow fastest)."""
    pytest.importorskip("pywr")
    import oracle.pywr_oracle_solver  # noqa: F401
    from benchmarks.workloads import build_model_param_sets
    from pywr_b200.device_run import run_on_device

    P, I = 3, 4
    host = build_model_param_sets(P, I, years=1, solver="oracle")
    host.run()
    dev = build_model_param_sets(P, I, years=1, solver="null")
    res = run_on_device(dev, engine_factory=OracleEngine)
    assert res.num_scenarios == P * I
    combos = [tuple(si.indices) for si in dev.scenarios.combinations]
    assert combos == [(p, i) for p in range(P) for i in range(I)]      # scenario indexing: bit-exact
    for name in ("volume1", "supplied1", "transferred", "deficit1"):
        a, b = dev.recorders[name], host.recorders[name]
        va = np.array(a.data) if hasattr(a, "data") else np.array(a.values())
        vb = np.array(b.data) if hasattr(b, "data") else np.array(b.values())
        assert rel_diff(va, vb) <= 1e-12, name
    # the parameter sets really differ: transferred volume per parameter set (mean over inflow)
    per_set = np.array(dev.recorders["transferred"].values()).reshape(P, I).mean(axis=1)
    assert np.ptp(per_set) > 0
