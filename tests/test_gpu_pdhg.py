"""GPU parity of the restarted-PDHG path (pywr_b200/csrc/pdhg_engine.cuh) through the C-ABI:
objective within 1e-7 relative of the simplex oracle (north_star tolerance), primal feasibility,
iteration counts against the numpy restatement, status codes, warm starts, options."""
import os

import numpy as np
import pytest

from fixtures_util import load_case, replay, synth_batch
from lp_numpy import residuals

pytestmark = pytest.mark.gpu
TOL = 1e-7


@pytest.fixture()
def force_pdhg(monkeypatch):
    monkeypatch.setenv("B200LP_FORCE_PDHG", "1")


def _gpu(s, S):
    from pywr_b200.engine import CudaEngine

    eng = CudaEngine(s, S, device=0)
    assert eng.stats()["kernel_variant"] == 200
    return eng


@pytest.mark.parametrize("case,S", [("network24.edge", 33), ("two_reservoir.edge", 64), ("thames.edge", 7)])
def test_pdhg_objective_and_feasibility(force_pdhg, case, S):
    from oracle.oracle_engine import OracleEngine

    s, tr = load_case(case)
    T = 5
    slots, days = synth_batch(s, tr, S, T, seed=11, common=True)
    gpu = _gpu(s, S)
    g = replay(gpu, s, slots, days)
    c = replay(OracleEngine(s, S), s, slots, days)
    assert g[3].sum() == 0 and c[3].sum() == 0
    err = np.abs(g[2] - c[2]) / np.maximum(1.0, np.abs(c[2]))
    assert err.max() < TOL, f"objective rel diff {err.max():.3e}"
    for t in range(T):
        for k in (0, S // 2, S - 1):
            viol, obj = residuals(s, slots[t], k, float(days[t]), g[1][t, :, k])
            assert viol < 1e-5 * max(1.0, np.abs(g[1][t, :, k]).max())
            assert abs(obj - g[2][t, k]) <= 1e-9 * max(1.0, abs(obj))
    # node flows are the column flows through the flow map (cython_glpk.pyx:1899-1911)
    F = np.zeros((s.n_nodes, s.n_cols))
    for v in range(s.n_nodes):
        for k in range(s.flow_indptr[v], s.flow_indptr[v + 1]):
            F[v, s.flow_col[k]] += s.flow_w[k]
    assert np.allclose(np.einsum("vj,tjs->tvs", F, g[1]), g[0], rtol=1e-12, atol=1e-9)
    st = gpu.stats()
    assert st["pdhg_iterations"] > 0 and 0 < st["reduced_rows"] < s.n_rows
    gpu.close()


def test_pdhg_iterations_match_numpy_restatement(force_pdhg):
    from oracle.pdhg_oracle import PdhgOracleEngine

    s, tr = load_case("network24.edge")
    S, T = 4, 3
    slots, days = synth_batch(s, tr, S, T, seed=3, common=True)
    gpu = _gpu(s, S)
    cpu = PdhgOracleEngine(s, S)
    g = replay(gpu, s, slots, days)
    c = replay(cpu, s, slots, days)
    assert np.max(np.abs(g[2] - c[2]) / np.maximum(1.0, np.abs(c[2]))) < TOL
    it_gpu, it_cpu = gpu.stats()["pdhg_iterations"], cpu.stats()["pdhg_iterations"]
    # same algorithm, same restart tests: the iteration totals agree up to rounding-driven restart shifts
    assert abs(it_gpu - it_cpu) <= 0.2 * it_cpu + 2 * cpu.check * S * T, (it_gpu, it_cpu)
    gpu.close()


def test_pdhg_nan_and_options(force_pdhg):
    s, tr = load_case("network24.edge")
    S = 5
    slots, days = synth_batch(s, tr, S, 2, seed=9, common=True)
    gpu = _gpu(s, S)
    gpu.set_option("pdhg_tol", 1e-5)
    a = replay(gpu, s, slots, days)
    assert a[3].sum() == 0
    loose = gpu.stats()["pdhg_iterations"]
    gpu.set_option("pdhg_tol", 1e-8)
    b = replay(gpu, s, slots, days)
    assert gpu.stats()["pdhg_iterations"] - loose > loose
    assert np.max(np.abs(a[2] - b[2]) / np.maximum(1.0, np.abs(b[2]))) < 1e-4
    bad = slots.copy()
    used = int(next(r for r in list(s.row_lb_ref) + list(s.row_ub_ref) if r >= 0))
    bad[:, used, 2] = np.nan
    c = replay(gpu, s, bad, days)
    assert (c[3] == 1).all() and gpu.status() == (2, 4)   # B200LP_ST_NAN, pywr GLPKError (cython_glpk.pyx:286-335)
    gpu.close()


def test_pdhg_warm_start(force_pdhg):
    s, tr = load_case("network24.edge")
    S = 8
    slots, days = synth_batch(s, tr, S, 1, seed=2, common=True)
    gpu = _gpu(s, S)
    sl = gpu.alloc_planes(s.n_slots)
    sl[:] = slots[0]
    ob = gpu.alloc_planes(1)
    gpu.reset(None)
    assert gpu.step(float(days[0]), sl, None, None, ob) == 0
    cold = gpu.stats()["pdhg_iterations"]
    first = ob.copy()
    assert gpu.step(float(days[0]), sl, None, None, ob) == 0
    warm = gpu.stats()["pdhg_iterations"] - cold
    assert warm <= 4 * 64 * S and warm * 4 < cold
    assert np.allclose(first, ob, rtol=1e-7)
    gpu.close()


def test_pdhg_tail_kernel_agrees_with_streaming_path(force_pdhg):
    """The CTA-per-scenario shared-memory tail kernel and the streaming kernels are the same algorithm:
    same objectives within tolerance, iteration totals within a restart or two."""
    s, tr = load_case("network24.edge")
    S, T = 40, 3
    slots, days = synth_batch(s, tr, S, T, seed=17, common=True)
    res, iters = [], []
    for tail in (0, 4096):
        gpu = _gpu(s, S)
        gpu.set_option("pdhg_tail", tail)
        res.append(replay(gpu, s, slots, days))
        iters.append(gpu.stats()["pdhg_iterations"])
        gpu.close()
    assert res[0][3].sum() == 0 and res[1][3].sum() == 0
    assert np.max(np.abs(res[0][2] - res[1][2]) / np.maximum(1.0, np.abs(res[1][2]))) < TOL
    assert abs(iters[0] - iters[1]) <= 0.25 * iters[0] + 64 * S * T, iters


def test_pdhg_on_the_2000_node_network_matches_highs(monkeypatch):
    """SURVEY 8(d) config 4 fixture (2,073 nodes, LP 3,994 x 2,672): no dictionary kernel fits, b200lp_create goes to
    the large-LP path; B200LP_LARGE_ENGINE=pdhg selects the PDHG iterations on it (the default there is the revised
    simplex, tests/test_gpu_rsx.py); scenario 0 replays the recorded planes and must reproduce HiGHS's optimum of
    every recorded timestep, scenario 1 is a world with half the inflow (no reference: convergence, and it
    cannot do better than scenario 0)."""
    from pywr_b200.engine import CudaEngine
    from pywr_b200.structure import LPStructure

    fix = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks", "fixtures", "network600.edge")
    s = LPStructure.load(fix + ".structure.npz")
    tr = np.load(fix + ".trace.npz")
    S, T = 2, 5
    monkeypatch.setenv("B200LP_LARGE_ENGINE", "pdhg")
    gpu = CudaEngine(s, S, device=0)
    st = gpu.stats()
    assert st["kernel_variant"] == 200 and st["reduced_rows"] < st["n_rows"] // 2
    vol_slots = sorted({int(r) for r in s.st_vol_ref if r >= 0})
    slots = gpu.alloc_planes(s.n_slots)
    cf = gpu.alloc_planes(s.n_cols)
    ob = gpu.alloc_planes(1)
    gpu.reset(None)
    for t in range(T):
        base = tr["slots"][t][:, 0]
        fin = np.isfinite(base)
        slots[:, 0] = base
        slots[:, 1] = np.where(fin, base * 0.5, base)
        slots[vol_slots, 1] = base[vol_slots]
        assert gpu.step(float(tr["days"][t]), slots, None, cf, ob) == 0
        ref = float(tr["objective_highs"][t, 0])
        assert abs(ob[0, 0] - ref) <= TOL * max(1.0, abs(ref)), (t, ob[0, 0], ref)
        assert ob[0, 1] >= ob[0, 0] - 1e-6 * abs(ob[0, 0])     # less water cannot give a better objective
        assert cf.min() >= 0.0 and cf[:, 1].sum() < cf[:, 0].sum()
    assert gpu.stats()["pdhg_iterations"] > 0
    gpu.close()


@pytest.mark.parametrize("case,S", [("network24.edge", 17), ("two_reservoir.edge", 5), ("network12.route", 9), ("thames.edge", 7)])
def test_pdhg_crossover_returns_a_vertex_with_a_valid_glpk_basis(force_pdhg, case, S):
    """north_star: "a batched restarted-PDHG path ... followed by crossover to a vertex".  After the first-order iterations
    every scenario's point is classified against its bounds, a basis is crashed from it and the revised simplex finishes
    (rsx_core.cuh::crossover).  The answer is then a VERTEX: objectives equal the dictionary oracle's to simplex accuracy
    (1e-9, not the first-order 1e-7), primal residuals at round-off, and b200lp_get_basis returns a valid GLPK-coded basis of
    the original LP (BasisManager, cython_glpk.pyx:197-232) on kernel variant 200.  With B200LP_PDHG_CROSSOVER=0 the path
    returns the first-order point and no basis, as before."""
    from lp_numpy import check_glpk_basis
    from oracle.oracle_engine import OracleEngine

    s, tr = load_case(case)
    slots, days = synth_batch(s, tr, S, 3, seed=13, common=True)
    T = len(days)
    gpu = _gpu(s, S)
    g = replay(gpu, s, slots, days)
    c = replay(OracleEngine(s, S), s, slots, days)
    assert g[3].sum() == 0 and c[3].sum() == 0
    err = np.abs(g[2] - c[2]) / np.maximum(1.0, np.abs(c[2]))
    assert err.max() < 1e-9, f"objective rel diff {err.max():.3e}"
    st = gpu.stats()
    assert st["pdhg_iterations"] > 0            # the iterations ran ...
    row, col = gpu.get_basis()                   # ... and a basis exists
    for k in (0, S // 2, S - 1):
        viol, obj = residuals(s, slots[T - 1], k, float(days[T - 1]), g[1][T - 1, :, k])
        assert viol < 1e-7 * max(1.0, np.abs(g[1][T - 1, :, k]).max())
        check_glpk_basis(s, slots[T - 1], float(days[T - 1]), k, g[1][T - 1, :, k], row[k], col[k])
    print(f"{case}: {st['pdhg_iterations']} PDHG scenario-iterations, {st['pivots']} simplex pivots after the crash "
          f"({st['pivots'] / (S * T):.1f} per scenario-timestep)")
    gpu.close()


def test_pdhg_without_crossover_has_no_basis(force_pdhg, monkeypatch):
    monkeypatch.setenv("B200LP_PDHG_CROSSOVER", "0")
    s, tr = load_case("thames.edge")
    slots, days = synth_batch(s, tr, 4, 2, seed=5, common=True)
    gpu = _gpu(s, 4)
    assert replay(gpu, s, slots, days)[3].sum() == 0
    with pytest.raises(RuntimeError):
        gpu.get_basis()
    gpu.close()
