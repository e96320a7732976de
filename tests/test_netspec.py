"""SURVEY.md 8(d) config 4 AS SPECIFIED, at test scale: the generated network with the specified mix — PiecewiseLinks,
AggregatedNodes with max_flow and with factors, a river tree plus transfers, per-scenario DAILY stochastic inflows
(benchmarks/workloads.py::spec_network_document / add_stochastic_inflows; tests/golden/make_netspec.py small: LP 373 x 243, 3
scenarios x 119 days).  The fixture's recorder arrays come from a host run of the reference framework whose LP solver is the
dictionary oracle; its traced timesteps carry HiGHS's optimum.  Checked here: the oracle's program mode, the host build of
the revised simplex, and on the GPU the large-LP path (revised simplex, and PDHG + crossover) through the C-ABI."""
import os

import numpy as np
import pytest

from fixtures_util import GOLDEN, rel_diff, replay, run_program
from pywr_b200.program import Program
from pywr_b200.structure import LPStructure

BASE = os.path.join(GOLDEN, "netspec", "small")


def _program():
    s = LPStructure.load(BASE + ".pstructure.npz")
    prog = Program.load(BASE + ".program.npz")
    z = np.load(BASE + ".expected.npz")
    return s, prog, z


def _trace():
    s = LPStructure.load(BASE + ".edge.structure.npz")
    tr = np.load(BASE + ".edge.trace.npz")
    return s, tr


def _check_program(prog, z, outs, tol):
    T = prog.n_steps
    names = [str(n) for n in z["names"]]
    assert len(names) == prog.n_recorders == 32
    worst = 0.0
    for i, (out, nm) in enumerate(zip(outs, names)):
        o = out / T if nm.startswith(("mf_", "df_")) else out
        d = rel_diff(o, z[f"rec{i}"])
        worst = max(worst, d)
        assert d <= tol, (nm, d)
    return worst


def test_structure_has_the_specified_mix():
    s, prog, z = _program()
    names = [str(n) for n in s.row_names]
    assert sum(n.startswith("factor:") for n in names) >= 1          # AggregatedNode with factors (ratio rows)
    assert sum(n.startswith("aggregated:") for n in names) >= 2      # AggregatedNode max_flow (licences)
    assert any("reach" in n and "Sublink" in n for n in [str(x) for x in s.node_names])   # PiecewiseLink sub-links
    assert prog.n_tables >= 4 and int(np.sum(prog.table_axis >= 0)) == 4               # 4 regional daily AR(1) series [T, S]
    assert len(s.dyn_a_nz) == 0


def test_oracle_program_equals_reference_host_run():
    from oracle.oracle_engine import OracleEngine

    s, prog, z = _program()
    eng = OracleEngine(s, 3)
    outs, status, _ = run_program(eng, prog)
    assert status[0] == 0
    _check_program(prog, z, outs, 1e-10)
    assert eng.counters()["pivots"] > 2 * 3 * prog.n_steps       # daily stochastic inflows: the basis changes most days
    eng.close()


def test_trace_objectives_equal_highs_on_both_cpu_simplexes():
    from oracle.oracle_engine import OracleEngine
    from oracle.rsx_oracle import RsxOracleEngine

    s, tr = _trace()
    S = tr["slots"].shape[2]
    for make in (OracleEngine, RsxOracleEngine):
        r = replay(make(s, S), s, tr["slots"], tr["days"])
        assert r[3].sum() == 0
        assert rel_diff(r[2], tr["objective_highs"]) <= 1e-9, make.__name__


@pytest.mark.gpu
def test_cuda_program_equals_reference_host_run():
    """b200lp_create picks the large-LP path (revised simplex, variant 300); parameters, Storage.after and recorders run as plane
    kernels around the batched solve.  north_star: flows and volumes within 1e-6 relative; asserted 1e-8."""
    from pywr_b200.engine import CudaEngine

    s, prog, z = _program()
    gpu = CudaEngine(s, 3, device=0)
    assert gpu.stats()["kernel_variant"] == 300
    outs, status, _ = run_program(gpu, prog, chunks=[10, 50])
    assert status[0] == 0
    worst = _check_program(prog, z, outs, 1e-8)
    st = gpu.stats()
    print(f"netspec small: worst relative difference vs the reference host run {worst:.2e}, {st['pivots']} pivots "
          f"({st['pivots'] / (3 * prog.n_steps):.1f} per scenario-timestep)")
    gpu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["rsx", "pdhg"])
def test_cuda_trace_objectives_equal_highs(monkeypatch, engine):
    """The traced timesteps through b200lp_step: the revised simplex, and the PDHG iterations followed by the crossover, both
    end at HiGHS's optimum (1e-9: vertices) with primal residuals at round-off."""
    from lp_numpy import residuals
    from pywr_b200.engine import CudaEngine

    monkeypatch.setenv("B200LP_LARGE_ENGINE", engine)
    s, tr = _trace()
    S = tr["slots"].shape[2]
    T = 6 if engine == "pdhg" else len(tr["days"])
    gpu = CudaEngine(s, S, device=0)
    assert gpu.stats()["kernel_variant"] == (300 if engine == "rsx" else 200)
    g = replay(gpu, s, tr["slots"][:T], tr["days"][:T])
    assert g[3].sum() == 0
    assert rel_diff(g[2], tr["objective_highs"][:T]) <= 1e-9
    for t in (0, T - 1):
        for k in range(S):
            viol, obj = residuals(s, tr["slots"][t], k, float(tr["days"][t]), g[1][t, :, k])
            assert viol < 1e-7 * max(1.0, np.abs(g[1][t, :, k]).max())
    gpu.close()


@pytest.mark.gpu
def test_cuda_full_scale_network_first_month_equals_reference_host_run():
    """The full-scale network of config 4 as specified (2,316 graph nodes: 150 reservoirs, 250 catchments, 301 outputs, 40
    aggregated nodes, 60 piecewise links; LP 3,105 x 2,010; benchmarks/fixtures/netspec2000*): the first 30 days of the
    device-resident run against the reference framework's host run (array recorders), and the traced timesteps against HiGHS."""
    from pywr_b200.engine import CudaEngine
    from pywr_b200.program import ARRAY_KINDS

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks", "fixtures")
    base = os.path.join(root, "netspec2000")
    s = LPStructure.load(base + ".pstructure.npz")
    prog = Program.load(base + ".program.npz")
    z = np.load(base + ".expected.npz")
    T = 30
    gpu = CudaEngine(s, 2, device=0)
    assert gpu.stats()["kernel_variant"] == 300
    gpu.load_program(prog)
    gpu.run(0, T)
    gpu.sync()
    assert gpu.program_status()[0] == 0
    checked = 0
    for r in range(prog.n_recorders):
        if int(prog.rec_kind[r]) in ARRAY_KINDS:
            out = gpu.fetch_recorder(r, prog.rec_shape(r, 2))
            assert rel_diff(out[:T], z[f"rec{r}"][:T]) <= 1e-8, str(z["names"][r])
            checked += 1
    assert checked >= 10
    gpu.close()
    st = LPStructure.load(base + ".edge.structure.npz")
    tr = np.load(base + ".edge.trace.npz")
    gpu = CudaEngine(st, tr["slots"].shape[2], device=0)
    g = replay(gpu, st, tr["slots"], tr["days"])
    assert g[3].sum() == 0 and rel_diff(g[2], tr["objective_highs"]) <= 1e-9
    gpu.close()


def test_full_scale_fixture_two_independent_simplexes_agree():
    """benchmarks/fixtures/netspec2000.expected.npz: host run of the reference framework over all 365 days with the DICTIONARY
    oracle as LP solver (15,121 pivots, 51 minutes); netspec2000_rsx.expected.npz: the same run with the host build of the
    revised simplex (13,144 pivots, 13 seconds).  Two independent simplex implementations, one set of recorder arrays."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks", "fixtures")
    a, b = np.load(os.path.join(root, "netspec2000.expected.npz")), np.load(os.path.join(root, "netspec2000_rsx.expected.npz"))
    assert [str(n) for n in a["names"]] == [str(n) for n in b["names"]]
    for i in range(len(a["names"])):
        assert rel_diff(a[f"rec{i}"], b[f"rec{i}"]) <= 1e-9, str(a["names"][i])
