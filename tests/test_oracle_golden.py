"""CPU tests (no GPU): the oracle against the committed golden vectors.

The golden traces were recorded from the reference framework (pywr built without GLPK) driving the
oracle through Pywr's solver API on the reference's own model documents
(tests/golden/make_golden.py); objective_highs is the optimum computed independently by scipy's
HiGHS for every recorded scenario-step.  The reference's known-answer suite itself is run against
the oracle with oracle/run_reference_tests.sh (592 passed)."""
import numpy as np
import pytest

from fixtures_util import golden_cases, load_case, replay, synth_batch
from lp_numpy import highs_objective, residuals
from oracle.oracle_engine import OracleEngine


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_replays_golden_bit_exact(name):
    s, tr = load_case(name)
    eng = OracleEngine(s, tr["slots"].shape[2])
    nf, cf, ob, nbad = replay(eng, s, tr["slots"], tr["days"])
    assert nbad.sum() == 0
    # bit-exact: same code, same inputs, same basis history
    assert np.array_equal(nf, tr["node_flow"])
    assert np.array_equal(cf, tr["col_flow"])
    assert np.array_equal(ob, tr["objective"])
    eng.close()


@pytest.mark.parametrize("name", golden_cases())
def test_golden_objective_matches_highs(name):
    s, tr = load_case(name)
    ok = np.isfinite(tr["objective_highs"])
    assert ok.all()
    # north_star: per-timestep objectives within 1e-7 relative
    np.testing.assert_allclose(tr["objective"], tr["objective_highs"], rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("name", ["two_reservoir.route", "thames.route", "thames.edge", "three_storage.route"])
def test_oracle_synthetic_batch_vs_highs(name):
    """Perturbed scenario batch: optimal objective equals HiGHS and the returned vertex is feasible."""
    s, tr = load_case(name)
    S, T = 16, 6
    slots, days = synth_batch(s, tr, S, T, seed=11)
    eng = OracleEngine(s, S)
    nf, cf, ob, nbad = replay(eng, s, slots, days)
    for t in range(T):
        for k in range(S):
            h = highs_objective(s, slots[t], k, days[t])
            if h is None:
                assert nbad[t] > 0
                continue
            assert abs(ob[t, k] - h) <= 1e-7 * max(1.0, abs(h))
            viol, obj = residuals(s, slots[t], k, days[t], cf[t, :, k])
            assert viol <= 1e-7 * max(1.0, np.abs(cf[t, :, k]).max())
            assert abs(obj - ob[t, k]) <= 1e-9 * max(1.0, abs(obj))
    eng.close()


def test_oracle_threads_give_identical_results():
    s, tr = load_case("two_reservoir.route")
    slots, days = synth_batch(s, tr, 64, 10, seed=3)
    a = OracleEngine(s, 64, threads=1)
    b = OracleEngine(s, 64, threads=4)
    ra = replay(a, s, slots, days)
    rb = replay(b, s, slots, days)
    for x, y in zip(ra, rb):
        assert np.array_equal(x, y)
    a.close(); b.close()


def test_oracle_infeasible_and_nan_status():
    s, tr = load_case("reservoir2.route")
    eng = OracleEngine(s, 1)
    slots = tr["slots"][0].copy()
    out = eng.alloc_planes(s.n_nodes)
    eng.reset(None)
    assert eng.step(1.0, slots, out, None, None) == 0
    bad = slots.copy(); bad[0, 0] = np.nan  # the volume plane
    assert eng.step(1.0, bad, out, None, None) == 1
    assert eng.status() == (0, 4)
    eng.close()
