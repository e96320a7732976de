"""The PDHG restatement (oracle/pdhg_oracle.py) against the golden traces: objective within 1e-7
relative of scipy HiGHS and of the simplex oracle, feasible within tolerance, warm starts cheaper
than cold starts.  CPU only."""
import numpy as np
import pytest

from fixtures_util import load_case, replay, synth_batch
from lp_numpy import residuals
from oracle.oracle_engine import OracleEngine
from oracle.pdhg_oracle import PdhgOracleEngine, PdhgPlan

TOL = 1e-7  # north_star: per-timestep objectives within 1e-7 relative (engine default KKT tolerance 1e-8)


def test_presolve_shapes():
    s, _ = load_case("network24.edge")
    p = PdhgPlan(s)
    # most rows of the edge form are single-column bounds or can never bind
    assert p.m_r < s.n_rows // 2 and p.m_r + len(p.sing_row) <= s.n_rows
    assert 0.5 < p.norm_a <= 1.0 + 1e-9   # Pock-Chambolle scaling bounds the 2-norm by 1
    assert np.all(p.dr > 0) and np.all(p.dc > 0)


@pytest.mark.parametrize("case", ["network24.edge", "two_reservoir.edge", "thames.edge"])
def test_pdhg_oracle_matches_highs(case):
    s, tr = load_case(case)
    T = min(8, len(tr["days"]))
    eng = PdhgOracleEngine(s, tr["slots"].shape[2])
    nf, cf, ob, nbad = replay(eng, s, tr["slots"][:T], tr["days"][:T])
    assert nbad.sum() == 0
    ref = tr["objective_highs"][:T]
    assert np.max(np.abs(ob - ref) / np.maximum(1.0, np.abs(ref))) < TOL
    for t in range(T):
        viol, _ = residuals(s, tr["slots"][t], 0, float(tr["days"][t]), cf[t, :, 0])
        assert viol < 1e-5 * max(1.0, np.abs(cf[t]).max())


def test_pdhg_oracle_batch_vs_simplex_oracle():
    s, tr = load_case("network24.edge")
    S, T = 3, 4
    slots, days = synth_batch(s, tr, S, T, seed=5, common=True)
    a = replay(PdhgOracleEngine(s, S), s, slots, days)
    b = replay(OracleEngine(s, S), s, slots, days)
    assert a[3].sum() == 0 and b[3].sum() == 0
    assert np.max(np.abs(a[2] - b[2]) / np.maximum(1.0, np.abs(b[2]))) < TOL


def test_warm_start_is_cheaper():
    s, tr = load_case("network24.edge")
    eng = PdhgOracleEngine(s, 1)
    replay(eng, s, tr["slots"][:1], tr["days"][:1])
    cold = int(eng.last_iterations[0])
    sl = eng.alloc_planes(s.n_slots)
    sl[:] = tr["slots"][0]
    eng.step(float(tr["days"][0]), sl, None, None, None)   # same LP again, warm
    assert int(eng.last_iterations[0]) <= 4 * eng.check < cold
