"""GPU tests (-m gpu) of the device-resident run on the LARGE-LP path (pywr_b200/csrc/plane_kernels.cuh,
plane_program.inc): the golden programs — parameters, Storage.after, virtual / rolling licences, recorders — executed as
plane kernels around the batched revised-simplex (or PDHG) solve of every timestep, against the oracle's program mode
and the recorder arrays of the reference framework's host runs.  Small LPs are routed to that path with
B200LP_FORCE_PDHG / B200LP_FORCE_RSX; the 2,073-node network takes it by itself."""
import numpy as np
import pytest

from fixtures_util import rel_diff, load_program_case, n_scenarios_of, program_cases, run_program, widen_program

pytestmark = pytest.mark.gpu


@pytest.fixture()
def force_rsx(monkeypatch):
    monkeypatch.setenv("B200LP_FORCE_PDHG", "1")
    monkeypatch.setenv("B200LP_FORCE_RSX", "1")


@pytest.fixture()
def force_pdhg(monkeypatch):
    monkeypatch.setenv("B200LP_FORCE_PDHG", "1")


def _engines(s, S, variant):
    from oracle.oracle_engine import OracleEngine
    from pywr_b200.engine import CudaEngine

    gpu = CudaEngine(s, S, device=0)
    assert gpu.stats()["kernel_variant"] == variant
    return OracleEngine(s, S, threads=8), gpu


def _static_programs():
    out = []
    for name in program_cases():
        s, _, _, _ = load_program_case(name)
        if len(s.dyn_a_nz) == 0:
            out.append(name)
    return out


@pytest.mark.parametrize("name", _static_programs())
def test_plane_program_matches_reference_host_run_and_oracle(force_rsx, name):
    s, prog, expected, names = load_program_case(name)
    cpu, gpu = _engines(s, n_scenarios_of(s, prog), 300)
    o_cpu, st_cpu, x_cpu = run_program(cpu, prog)
    o_gpu, st_gpu, x_gpu = run_program(gpu, prog)
    assert st_gpu == st_cpu and st_gpu[0] == 0
    T = prog.n_steps
    for out_g, out_c, exp, nm in zip(o_gpu, o_cpu, expected, names):
        assert rel_diff(out_g, out_c) <= 1e-9, (nm, rel_diff(out_g, out_c))
        if nm.startswith(("mf_", "df_")):
            out_g = out_g / T
        assert rel_diff(out_g, exp) <= 1e-9, (nm, rel_diff(out_g, exp))   # north_star asks 1e-6 of the reference run
    assert rel_diff(x_gpu[1], x_cpu[1]) <= 1e-9 and rel_diff(x_gpu[2], x_cpu[2]) <= 1e-9   # final volumes, pc
    gpu.close(); cpu.close()


@pytest.mark.parametrize("name,S,seed", [("two_reservoir.route", 257, 1), ("thames.edge", 333, 3), ("demand_saving2.route", 64, 4)])
def test_plane_program_widened_parity_and_chunks(force_rsx, name, S, seed):
    s, prog, _, names = load_program_case(name)
    q = widen_program(s, prog, S, seed)
    cpu, gpu = _engines(s, S, 300)
    o_cpu, st_cpu, x_cpu = run_program(cpu, q)
    o_gpu, st_gpu, x_gpu = run_program(gpu, q, chunks=[50, 200])
    assert st_gpu == st_cpu
    for out_g, out_c, nm in zip(o_gpu, o_cpu, names):
        assert rel_diff(out_g, out_c) <= 1e-9, (nm, rel_diff(out_g, out_c))
    assert rel_diff(x_gpu[1], x_cpu[1]) <= 1e-9
    # same answers after a program reset (Model.reset): fresh seed, fresh bases
    gpu.program_reset()
    gpu.run(0, q.n_steps); gpu.sync()
    again = [gpu.fetch_recorder(r, q.rec_shape(r, S)) for r in range(q.n_recorders)]
    for a, b in zip(o_gpu, again):
        assert np.array_equal(a, b, equal_nan=True)
    gpu.close(); cpu.close()


def test_plane_program_on_the_pdhg_iterations(force_pdhg):
    """the same plane kernels around the first-order method: recorders within its tolerance of the oracle's"""
    s, prog, _, names = load_program_case("thames.route")
    S = 8
    q = widen_program(s, prog, S, seed=5)
    q.n_steps = min(q.n_steps, 40)
    cpu, gpu = _engines(s, S, 200)
    o_cpu, st_cpu, _ = run_program(cpu, q)
    o_gpu, st_gpu, _ = run_program(gpu, q)
    assert st_gpu == st_cpu and st_gpu[0] == 0
    for out_g, out_c, nm in zip(o_gpu, o_cpu, names):
        assert rel_diff(out_g, out_c) <= 1e-5, (nm, rel_diff(out_g, out_c))
    gpu.close(); cpu.close()


def test_plane_program_failed_scenario_stops_and_is_reported(force_rsx):
    """a NaN inflow at timestep 7 of scenario 3: that scenario stops there (status NaN, GLPKError in the reference,
    cython_glpk.pyx:286-335), the others are untouched — same report as the oracle"""
    s, prog, _, names = load_program_case("two_reservoir.route")
    S = 6
    q = widen_program(s, prog, S, seed=2)
    k = next(i for i in range(q.n_tables) if q.table_rows[i] == q.n_steps and q.table_cols[i] >= S)
    data = q.table_data.copy()
    data[q.table_offset[k] + 7 * q.table_cols[k] + 3] = np.nan
    q.table_data = data
    cpu, gpu = _engines(s, S, 300)
    o_cpu, st_cpu, _ = run_program(cpu, q)
    o_gpu, st_gpu, _ = run_program(gpu, q)
    assert st_gpu == st_cpu and st_gpu[0] == 1 and st_gpu[1] == 3 and st_gpu[3] == 7
    ok = [i for i in range(S) if i != 3]
    for out_g, out_c, nm in zip(o_gpu, o_cpu, names):
        a = out_g[..., ok] if out_g.ndim > 1 else out_g[ok]
        b = out_c[..., ok] if out_c.ndim > 1 else out_c[ok]
        assert rel_diff(a, b) <= 1e-9, nm
    gpu.close(); cpu.close()


def test_plane_program_pipelined_equals_serial(force_rsx):
    s, prog, _, _ = load_program_case("thames.route")
    S = 64
    q = widen_program(s, prog, S, seed=21)
    cpu, gpu = _engines(s, S, 300)
    serial, st, _ = run_program(gpu, q)
    tables = {}
    for k in range(q.n_tables):
        r0, c0 = int(q.table_rows[k]), int(q.table_cols[k])
        a = gpu.alloc_pinned((r0, c0))
        a[...] = q.table_data[q.table_offset[k]: q.table_offset[k] + r0 * c0].reshape(r0, c0)
        tables[k] = a
        gpu.update_table(k, np.zeros((r0, c0)))
    gpu.sync()
    outs = [gpu.alloc_pinned(q.rec_shape(r, S)) for r in range(q.n_recorders)]
    gpu.run_pipelined(tables, outs, chunks=8)
    assert gpu.program_status() == st
    for a, b in zip(serial, outs):
        assert np.array_equal(a, b, equal_nan=True)
    gpu.close(); cpu.close()


def test_plane_program_on_the_2000_node_network_matches_the_reference_host_run():
    """SURVEY 8(d) config 4 end to end on the device: the 2,073-node network (LP 3,994 x 2,672, 1,201 ops, 601 tables, 30
    recorders) for a year of daily timesteps.  b200lp_create picks the revised simplex by itself; the recorder arrays must
    equal those of the reference framework's HOST run (Model.run with the reference's own Parameter / Storage.after /
    Recorder code, tests/golden/make_network_program.py) within north_star's 1e-6 — observed: round-off."""
    import os

    from pywr_b200.engine import CudaEngine
    from pywr_b200.program import Program
    from pywr_b200.structure import LPStructure

    fix = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks", "fixtures", "network600")
    s = LPStructure.load(fix + ".pstructure.npz")
    prog = Program.load(fix + ".program.npz")
    z = np.load(fix + ".expected.npz")
    zd = np.load(fix + ".expected_dict.npz")    # the same program on the dictionary oracle (independent simplex)
    names = [str(n) for n in z["names"]]
    S = n_scenarios_of(s, prog)
    gpu = CudaEngine(s, S, device=0)
    assert gpu.stats()["kernel_variant"] == 300
    outs, st, state = run_program(gpu, prog)
    assert st[0] == 0
    T = prog.n_steps
    worst = 0.0
    for i, (o, nm) in enumerate(zip(outs, names)):
        if nm.startswith(("mf_", "df_")):
            o = o / T
        d = rel_diff(o, z[f"rec{i}"])
        worst = max(worst, d)
        assert d <= 1e-6, (nm, d)
        od = zd[f"rec{i}"] / T if nm.startswith(("mf_", "df_")) else zd[f"rec{i}"]
        assert rel_diff(o, od) <= 1e-6, (nm, rel_diff(o, od))
    stt = gpu.stats()
    assert stt["refactors"] == 0 and stt["pivots"] > 1000
    print(f"network600: T={T} S={S}, {len(names)} recorders, worst relative difference vs the reference host run {worst:.2e}, "
          f"{stt['pivots']} pivots")
    gpu.close()
