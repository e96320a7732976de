"""GPU parity of the revised-simplex method of the large-LP path (pywr_b200/csrc/rsx_core.cuh, rsx_kernels.cuh)
through the C-ABI: objectives equal to the dictionary oracle's (a vertex method: 1e-9, far inside north_star's 1e-7),
primal feasibility, the same pivot counts as the host build of the same core, status codes, warm starts, reset,
and the 2,073-node network against HiGHS."""
import os

import numpy as np
import pytest

from fixtures_util import load_case, replay, synth_batch
from lp_numpy import check_glpk_basis, residuals

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture()
def force_rsx(monkeypatch):
    monkeypatch.setenv("B200LP_FORCE_PDHG", "1")   # route small LPs to the large-LP front end ...
    monkeypatch.setenv("B200LP_FORCE_RSX", "1")    # ... and solve them with the revised simplex


def _gpu(s, S):
    from pywr_b200.engine import CudaEngine

    eng = CudaEngine(s, S, device=0)
    assert eng.stats()["kernel_variant"] == 300
    return eng


@pytest.mark.parametrize("case,S", [("network24.edge", 333), ("two_reservoir.edge", 64), ("thames.edge", 7), ("network12.route", 40),
                                    ("piecewise1.edge", 5), ("aggregated1.route", 16)])
def test_rsx_objective_feasibility_and_pivots(force_rsx, case, S):
    from oracle.oracle_engine import OracleEngine
    from oracle.rsx_oracle import RsxOracleEngine

    s, tr = load_case(case)
    slots, days = synth_batch(s, tr, S, 5, seed=11, common=True)
    T = len(days)
    gpu = _gpu(s, S)
    g = replay(gpu, s, slots, days)
    c = replay(OracleEngine(s, S), s, slots, days)
    assert g[3].sum() == 0 and c[3].sum() == 0
    err = np.abs(g[2] - c[2]) / np.maximum(1.0, np.abs(c[2]))
    assert err.max() < TOL, f"objective rel diff {err.max():.3e}"
    for t in range(T):
        for k in (0, S // 2, S - 1):
            viol, obj = residuals(s, slots[t], k, float(days[t]), g[1][t, :, k])
            assert viol < 1e-7 * max(1.0, np.abs(g[1][t, :, k]).max())
            assert abs(obj - g[2][t, k]) <= 1e-9 * max(1.0, abs(obj))
    F = np.zeros((s.n_nodes, s.n_cols))
    for v in range(s.n_nodes):
        for k in range(s.flow_indptr[v], s.flow_indptr[v + 1]):
            F[v, s.flow_col[k]] += s.flow_w[k]
    assert np.allclose(np.einsum("vj,tjs->tvs", F, g[1]), g[0], rtol=1e-12, atol=1e-9)
    # the host build of the same core takes the same pivots (reductions differ only in summation order)
    h = RsxOracleEngine(s, S)
    hres = replay(h, s, slots, days)
    st = gpu.stats()
    assert abs(st["pivots"] - h.stats()["pivots"]) <= 0.02 * h.stats()["pivots"] + 2, (st["pivots"], h.stats()["pivots"])
    assert np.max(np.abs(hres[2] - g[2]) / np.maximum(1.0, np.abs(g[2]))) < TOL
    assert st["refactors"] == 0 and 0 <= st["reduced_rows"] < s.n_rows
    gpu.close()


def test_rsx_nan_status_warm_start_and_reset(force_rsx):
    s, tr = load_case("network24.edge")
    S = 6
    slots, days = synth_batch(s, tr, S, 2, seed=9, common=True)
    gpu = _gpu(s, S)
    a = replay(gpu, s, slots, days)
    assert a[3].sum() == 0
    cold = gpu.stats()["pivots"]
    sl = gpu.alloc_planes(s.n_slots)
    sl[:] = slots[1]
    ob = gpu.alloc_planes(1)
    assert gpu.step(float(days[1]), sl, None, None, ob) == 0      # same LP again: every stored basis is optimal
    assert gpu.stats()["pivots"] == cold and np.allclose(ob[0], a[2][1], rtol=1e-12)
    b = replay(gpu, s, slots, days)                                # replay() resets: same answers from a fresh start
    assert np.array_equal(a[2], b[2])
    bad = slots.copy()
    used = int(next(r for r in list(s.row_lb_ref) + list(s.row_ub_ref) if r >= 0))
    bad[:, used, 2] = np.nan
    c = replay(gpu, s, bad, days)
    assert (c[3] == 1).all() and gpu.status() == (2, 4)   # B200LP_ST_NAN, pywr GLPKError (cython_glpk.pyx:286-335)
    ok = [k for k in range(S) if k != 2]
    assert np.allclose(c[2][:, ok], a[2][:, ok], rtol=1e-12)
    gpu.close()


def test_rsx_infeasible_scenario_is_reported(force_rsx):
    """A catchment that must deliver a NEGATIVE flow (lb = ub = -5 on a row whose columns are all x >= 0 with coefficient +1)
    has no feasible solution: status 1 ("no primal feasible solution", cython_glpk.pyx:1060) for that scenario only, exactly
    like the dictionary oracle; the other scenarios are solved as if nothing had happened."""
    from oracle.oracle_engine import OracleEngine

    s, tr = load_case("network24.edge")
    S = 4
    slots, days = synth_batch(s, tr, S, 1, seed=3, common=True)
    i = next(k for k, nm in enumerate(s.row_names) if str(nm).startswith("flow:catch") and s.row_lb_ref[k] >= 0 and s.row_ub_ref[k] >= 0)
    gpu = _gpu(s, S)
    good = replay(gpu, s, slots, days)
    assert good[3].sum() == 0
    bad = slots.copy()
    bad[:, int(s.row_lb_ref[i]), 1] = -5.0      # min_flow = max_flow = -5 for scenario 1
    bad[:, int(s.row_ub_ref[i]), 1] = -5.0
    c = replay(gpu, s, bad, days)
    assert c[3].sum() == 1                     # the step reports failed scenarios ...
    assert gpu.status() == (1, 1)              # ... the first (only) one is scenario 1, B200LP_ST_INFEASIBLE
    o = replay(OracleEngine(s, S), s, bad, days)
    assert o[3].sum() == 1
    ok = [0, 2, 3]
    assert np.allclose(c[2][:, ok], good[2][:, ok], rtol=1e-12)
    gpu.close()


def test_rsx_is_selected_for_the_2000_node_network_and_matches_highs():
    """SURVEY 8(d) config 4 fixture (2,073 nodes, LP 3,994 x 2,672): no dictionary kernel fits; b200lp_create picks the
    revised simplex (variant 300) by itself.  Scenario 0 replays the recorded planes and must reproduce HiGHS's optimum
    of every recorded timestep; the other scenarios are worlds with less inflow."""
    from pywr_b200.engine import CudaEngine
    from pywr_b200.structure import LPStructure

    fix = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks", "fixtures", "network600.edge")
    s = LPStructure.load(fix + ".structure.npz")
    tr = np.load(fix + ".trace.npz")
    S, T = 24, 6
    gpu = CudaEngine(s, S, device=0)
    st = gpu.stats()
    assert st["kernel_variant"] == 300 and st["reduced_rows"] < st["n_rows"] // 2
    vol_slots = sorted({int(r) for r in s.st_vol_ref if r >= 0})
    slots = gpu.alloc_planes(s.n_slots)
    cf = gpu.alloc_planes(s.n_cols)
    ob = gpu.alloc_planes(1)
    gpu.reset(None)
    scale = np.linspace(1.0, 0.4, S)
    for t in range(T):
        base = tr["slots"][t][:, 0]
        fin = np.isfinite(base)
        slots[:] = np.where(fin, base, 0.0)[:, None] * scale[None, :]
        slots[~fin] = base[~fin][:, None]
        slots[vol_slots] = base[vol_slots][:, None]
        assert gpu.step(float(tr["days"][t]), slots, None, cf, ob) == 0
        ref = float(tr["objective_highs"][t, 0])
        assert abs(ob[0, 0] - ref) <= TOL * max(1.0, abs(ref)), (t, ob[0, 0], ref)
        assert np.all(np.diff(ob[0]) >= -1e-6 * abs(ob[0, 0]))     # less water cannot give a better objective
        assert cf.min() >= 0.0
        for k in (0, S - 1):
            viol, obj = residuals(s, slots, k, float(tr["days"][t]), cf[:, k])
            assert viol < 1e-6 * max(1.0, np.abs(cf[:, k]).max())
            assert abs(obj - ob[0, k]) <= 1e-9 * max(1.0, abs(obj))
    st = gpu.stats()
    assert st["pivots"] > 500 and st["refactors"] == 0
    gpu.close()


@pytest.mark.parametrize("case,S", [("network24.edge", 17), ("two_reservoir.edge", 5), ("network12.route", 9)])
def test_rsx_get_basis_is_a_valid_basis_of_the_original_lp(force_rsx, case, S):
    """b200lp_get_basis on the large-LP path (BasisManager.row_stat / col_stat, cython_glpk.pyx:197-232): GLPK-coded statuses
    in the ORIGINAL row space.  Exactly m variables are basic, every non-basic variable sits on the bound its status names,
    and the basic columns / rows form a non-singular basis matrix whose solution is the engine's x."""
    s, tr = load_case(case)
    slots, days = synth_batch(s, tr, S, 3, seed=13, common=True)
    T = len(days)
    gpu = _gpu(s, S)
    g = replay(gpu, s, slots, days)
    assert g[3].sum() == 0
    row, col = gpu.get_basis()
    for k in (0, S // 2, S - 1):
        check_glpk_basis(s, slots[T - 1], float(days[T - 1]), k, g[1][T - 1, :, k], row[k], col[k])
    gpu.close()
