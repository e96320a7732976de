"""The revised-simplex core (pywr_b200/csrc/rsx_core.cuh) compiled for the host (oracle/rsx_host.cpp, one-thread
context) against the golden traces: objectives equal to scipy HiGHS's and the dictionary oracle's, feasible
points, warm starts that need no pivots, the seeded first timestep.  CPU only: this is how the algorithm the
CUDA library runs is checked where no GPU exists; tests/test_gpu_rsx.py checks the device build."""
import os

import numpy as np
import pytest

from fixtures_util import golden_cases, load_case, replay, synth_batch
from lp_numpy import residuals
from oracle.oracle_engine import OracleEngine
from oracle.rsx_oracle import RsxOracleEngine

TOL = 1e-9  # a vertex method: far inside north_star's 1e-7


def _static_cases():
    out = []
    for name in golden_cases():
        s, _ = load_case(name)
        if len(s.dyn_a_nz) == 0:
            out.append(name)
    return out


@pytest.mark.parametrize("case", _static_cases())
def test_rsx_core_matches_highs_on_golden_traces(case):
    s, tr = load_case(case)
    T = min(10, len(tr["days"]))
    eng = RsxOracleEngine(s, tr["slots"].shape[2])
    nf, cf, ob, nbad = replay(eng, s, tr["slots"][:T], tr["days"][:T])
    assert nbad.sum() == 0
    ref = tr["objective_highs"][:T]
    assert np.max(np.abs(ob - ref) / np.maximum(1.0, np.abs(ref))) < TOL
    for t in range(T):
        viol, obj = residuals(s, tr["slots"][t], 0, float(tr["days"][t]), cf[t, :, 0])
        assert viol < 1e-7 * max(1.0, np.abs(cf[t]).max())
        assert abs(obj - ob[t, 0]) <= 1e-9 * max(1.0, abs(obj))


@pytest.mark.parametrize("case,S", [("network24.edge", 9), ("two_reservoir.edge", 8), ("thames.edge", 7), ("network12.route", 6)])
def test_rsx_core_batch_vs_dictionary_oracle(case, S):
    s, tr = load_case(case)
    T = 6
    slots, days = synth_batch(s, tr, S, T, seed=11, common=True)
    eng = RsxOracleEngine(s, S)
    a = replay(eng, s, slots, days)
    b = replay(OracleEngine(s, S), s, slots, days)
    assert a[3].sum() == 0 and b[3].sum() == 0
    assert np.max(np.abs(a[2] - b[2]) / np.maximum(1.0, np.abs(b[2]))) < TOL
    st = eng.stats()
    assert st["pivots"] > 0 and st["refactors"] == 0


def test_rsx_warm_start_needs_no_pivots():
    s, tr = load_case("network24.edge")
    eng = RsxOracleEngine(s, 1)
    replay(eng, s, tr["slots"][:1], tr["days"][:1])
    cold = eng.stats()["pivots"]
    sl = eng.alloc_planes(s.n_slots)
    sl[:] = tr["slots"][0]
    ob = eng.alloc_planes(1)
    assert eng.step(float(tr["days"][0]), sl, None, None, ob) == 0   # same LP again: the stored basis is optimal
    assert cold > 0 and eng.stats()["pivots"] == cold


def test_rsx_core_on_the_2000_node_network():
    """config 4 fixture (LP 3,994 x 2,672, reduced 1,242 x 2,672): scenario 0 reproduces HiGHS's optimum of every recorded
    timestep; the seeded scenarios need a few pivots per timestep, not a cold start each."""
    from pywr_b200.structure import LPStructure

    fix = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks", "fixtures", "network600.edge")
    s = LPStructure.load(fix + ".structure.npz")
    tr = np.load(fix + ".trace.npz")
    S, T = 2, 6
    eng = RsxOracleEngine(s, S)
    vol_slots = sorted({int(r) for r in s.st_vol_ref if r >= 0})
    slots = eng.alloc_planes(s.n_slots); cf = eng.alloc_planes(s.n_cols); ob = eng.alloc_planes(1)
    eng.reset(None)
    per_step = []
    for t in range(T):
        base = tr["slots"][t][:, 0]
        fin = np.isfinite(base)
        slots[:, 0] = base
        slots[:, 1] = np.where(fin, base * 0.5, base)
        slots[vol_slots, 1] = base[vol_slots]
        p0 = eng.stats()["pivots"]
        assert eng.step(float(tr["days"][t]), slots, None, cf, ob) == 0
        per_step.append(eng.stats()["pivots"] - p0)
        ref = float(tr["objective_highs"][t, 0])
        assert abs(ob[0, 0] - ref) <= TOL * max(1.0, abs(ref)), (t, ob[0, 0], ref)
        assert ob[0, 1] >= ob[0, 0] - 1e-6 * abs(ob[0, 0])
        assert cf.min() >= 0.0
    assert per_step[0] > 300 and max(per_step[1:]) < per_step[0] // 4, per_step
    assert eng.stats()["refactors"] == 0


def test_rsx_scratch_placement_takes_the_same_pivots():
    """mode 1 (reduced costs / pivot row / index lists outside the block's fast memory: two blocks per SM on the device) is a
    placement, not an algorithm: identical objectives and pivot counts"""
    s, tr = load_case("network24.edge")
    S, T = 5, 6
    slots, days = synth_batch(s, tr, S, T, seed=4, common=True)
    a_eng, b_eng = RsxOracleEngine(s, S, mode=0), RsxOracleEngine(s, S, mode=1)
    a = replay(a_eng, s, slots, days)
    b = replay(b_eng, s, slots, days)
    assert np.array_equal(a[2], b[2]) and a_eng.stats()["pivots"] == b_eng.stats()["pivots"] > 0


def test_rsx_drift_guard_refactorises_and_recovers():
    """A damaged basis inverse (here: noise written into W between two timesteps) is caught — by the pivot-element consistency
    test or by the residual check |A x - r| <= 1e-7 — and the basis is re-factorised from the slack basis by Gauss-Jordan pivots
    (rsx_core.cuh::refactor); the answers stay those of the dictionary oracle."""
    s, tr = load_case("network24.edge")
    S, T = 3, 6
    slots, days = synth_batch(s, tr, S, T, seed=8, common=True)
    ref = replay(OracleEngine(s, S), s, slots, days)
    eng = RsxOracleEngine(s, S)
    sl = eng.alloc_planes(s.n_slots); ob = eng.alloc_planes(1)
    eng.reset(None)
    rng = np.random.default_rng(0)
    for t in range(T):
        if t == 3:
            eng.W[1] += 1e-3 * rng.standard_normal(eng.W[1].shape)   # scenario 1 only
        sl[:] = slots[t]
        assert eng.step(float(days[t]), sl, None, None, ob) == 0
        assert np.max(np.abs(ob[0] - ref[2][t]) / np.maximum(1.0, np.abs(ref[2][t]))) < TOL, t
    assert eng.stats()["refactors"] >= 1


def test_rsx_scenarios_start_cold_when_the_seed_scenario_fails():
    """scenario 0 carries a NaN in the first timestep: nothing seeds the others, each starts from its own slack basis"""
    s, tr = load_case("network24.edge")
    S, T = 4, 3
    slots, days = synth_batch(s, tr, S, T, seed=6, common=True)
    used = int(next(r for r in list(s.row_lb_ref) + list(s.row_ub_ref) if r >= 0))
    bad = slots.copy()
    bad[0, used, 0] = np.nan
    a = replay(RsxOracleEngine(s, S), s, bad, days)
    b = replay(OracleEngine(s, S), s, slots, days)
    assert a[3][0] == 1 and a[3][1:].sum() == 0
    assert np.max(np.abs(a[2][:, 1:] - b[2][:, 1:]) / np.maximum(1.0, np.abs(b[2][:, 1:]))) < TOL
    assert np.max(np.abs(a[2][1:, 0] - b[2][1:, 0]) / np.maximum(1.0, np.abs(b[2][1:, 0]))) < TOL   # scenario 0 recovers next step


def test_rsx_reduced_cost_refresh_changes_nothing():
    """rebuilding the reduced costs from W and the costs every few pivots (the device does it every 1,000 pivots of a scenario to
    bound accumulated round-off) gives the same pivots and objectives as carrying them"""
    s, tr = load_case("network24.edge")
    S, T = 4, 10
    slots, days = synth_batch(s, tr, S, T, seed=12, common=True)
    a_eng, b_eng = RsxOracleEngine(s, S, refresh_pivots=0), RsxOracleEngine(s, S, refresh_pivots=3)
    a = replay(a_eng, s, slots, days)
    b = replay(b_eng, s, slots, days)
    assert a[3].sum() == 0 and b[3].sum() == 0
    assert np.max(np.abs(a[2] - b[2]) / np.maximum(1.0, np.abs(a[2]))) < 1e-12
    assert a_eng.stats()["pivots"] == b_eng.stats()["pivots"] > 3 * S


def test_config4_fixture_is_pinned_by_the_dictionary_oracle():
    """benchmarks/fixtures/network600.expected.npz (what the CUDA run of config 4 is compared with) comes from a host run
    of the reference framework whose LP solver is the host build of the product's own revised simplex.  The same device
    program executed by the DICTIONARY oracle's program mode — an independent dense simplex, with its own parameter /
    Storage.after / recorder restatement — over all 365 timesteps (tests/golden/make_network_program_dict.py, 5,136 pivots)
    gives the same recorder arrays: the fixture is not self-referential."""
    fix = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks", "fixtures", "network600")
    a, b = np.load(fix + ".expected.npz"), np.load(fix + ".expected_dict.npz")
    names = [str(n) for n in a["names"]]
    assert names == [str(n) for n in b["names"]] and len(names) == 30
    T = 365
    worst = 0.0
    for i, nm in enumerate(names):
        x, y = a[f"rec{i}"], b[f"rec{i}"]
        if nm.startswith(("mf_", "df_")):       # the reference divides by the number of timesteps in finish()
            y = y / T
        assert x.shape == y.shape
        d = float(np.max(np.abs(x - y) / np.maximum(1.0, np.abs(y))))
        worst = max(worst, d)
    assert worst <= 1e-9, worst
