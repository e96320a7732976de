"""GPU tests (-m gpu) of the device-resident run: CUDA run kernel through the C-ABI against the
oracle's program mode and against the recorder arrays of the reference framework's host run."""
import numpy as np
import pytest

from fixtures_util import rel_diff, load_program_case, n_scenarios_of, program_cases, run_program, widen_program

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return rel_diff(a, b)


def _engines(s, S):
    from oracle.oracle_engine import OracleEngine
    from pywr_b200.engine import CudaEngine

    return OracleEngine(s, S, threads=8), CudaEngine(s, S, device=0)


@pytest.mark.parametrize("name", program_cases())
def test_program_matches_reference_host_run_and_oracle(name):
    s, prog, expected, names = load_program_case(name)
    cpu, gpu = _engines(s, n_scenarios_of(s, prog))
    o_cpu, st_cpu, x_cpu = run_program(cpu, prog)
    o_gpu, st_gpu, x_gpu = run_program(gpu, prog)
    assert st_gpu == st_cpu and st_gpu[0] == 0
    T = prog.n_steps
    for out_g, out_c, exp, nm in zip(o_gpu, o_cpu, expected, names):
        assert _rel(out_g, out_c) <= 1e-9, (nm, _rel(out_g, out_c))
        if nm.startswith(("mf_", "df_")):
            out_g = out_g / T
        # north_star: flows and volumes within 1e-6 relative of the reference run
        assert _rel(out_g, exp) <= 1e-9, (nm, _rel(out_g, exp))
    for a, b in zip(x_gpu, x_cpu):
        assert _rel(a, b) <= 1e-9
    gpu.close(); cpu.close()


@pytest.mark.parametrize("name,S,seed", [("two_reservoir.route", 1024, 1), ("thames.route", 2048, 2),
                                          ("thames.edge", 333, 3), ("demand_saving2.route", 64, 4)])
def test_widened_program_parity(name, S, seed):
    """Scenario batches at benchmark-like sizes: every recorder, the final volumes, the last column
    flows and the pivot count agree with the oracle."""
    s, prog, _, names = load_program_case(name)
    q = widen_program(s, prog, S, seed)
    cpu, gpu = _engines(s, S)
    o_cpu, st_cpu, x_cpu = run_program(cpu, q)
    o_gpu, st_gpu, x_gpu = run_program(gpu, q, chunks=[50, 200])
    assert st_gpu == st_cpu
    worst = 0.0
    for out_g, out_c, nm in zip(o_gpu, o_cpu, names):
        worst = max(worst, _rel(out_g, out_c))
        assert _rel(out_g, out_c) <= 1e-9, (nm, _rel(out_g, out_c))
    for a, b in zip(x_gpu, x_cpu):
        assert _rel(a, b) <= 1e-9
    if st_cpu[0] == 0:
        assert gpu.counters()["pivots"] == cpu.counters()["pivots"]
    print(f"{name} S={S}: worst relative difference over {len(names)} recorders = {worst:.2e}")
    gpu.close(); cpu.close()


def test_program_reset_reproduces():
    s, prog, _, _ = load_program_case("thames.route")
    cpu, gpu = _engines(s, n_scenarios_of(s, prog))
    a, _, _ = run_program(gpu, prog)
    gpu.program_reset()
    gpu.run(0, prog.n_steps); gpu.sync()
    b = [gpu.fetch_recorder(r, prog.rec_shape(r, gpu.S)) for r in range(prog.n_recorders)]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    gpu.close(); cpu.close()


@pytest.mark.parametrize("name,S,chunks", [("two_reservoir.route", 257, 7), ("thames.route", 64, 16), ("thames.route", 5, 1000)])
def test_pipelined_run_equals_serial_run(name, S, chunks):
    """b200lp_run_pipelined (time slices, copies overlapped on three streams, host tables in, host
    recorders out) returns bit-identical recorder arrays to load + run + fetch."""
    s, prog, _, _ = load_program_case(name)
    q = widen_program(s, prog, S, seed=21)
    cpu, gpu = _engines(s, S)
    serial, st, _ = run_program(gpu, q)
    tables = {}
    for k in range(q.n_tables):
        r0, c0 = int(q.table_rows[k]), int(q.table_cols[k])
        a = gpu.alloc_pinned((r0, c0))
        a[...] = q.table_data[q.table_offset[k]: q.table_offset[k] + r0 * c0].reshape(r0, c0)
        tables[k] = a
        gpu.update_table(k, np.zeros((r0, c0)))     # the pipelined call must bring the real contents itself
    gpu.sync()
    outs = [gpu.alloc_pinned(q.rec_shape(r, S)) for r in range(q.n_recorders)]
    gpu.run_pipelined(tables, outs, chunks=chunks)
    assert gpu.program_status() == st
    for a, b in zip(serial, outs):
        assert np.array_equal(a, b, equal_nan=True)
    gpu.close(); cpu.close()


@pytest.mark.parametrize("name,S,seed", [("two_reservoir.route", 300, 5), ("thames.route", 1000, 6), ("demand_saving2.route", 64, 7)])
def test_run_kernel_register_builds_agree(monkeypatch, name, S, seed):
    """The run kernel exists in a 168-register build (3 blocks per SM: shortest chain) and a 128-register build (4 blocks
    per SM: a third more resident warps; chosen when the batch would otherwise need another wave, b200lp.cu::pick_run_blocks).
    Same source, same arithmetic: bit-identical recorders, states and pivot counts, both equal to the oracle's."""
    s, prog, _, names = load_program_case(name)
    q = widen_program(s, prog, S, seed)
    res = {}
    for blocks in ("3", "4"):
        monkeypatch.setenv("B200LP_RUN_BLOCKS", blocks)
        cpu, gpu = _engines(s, S)
        res[blocks] = run_program(gpu, q) + (gpu.counters()["pivots"],)
        if blocks == "4":
            o_cpu, st_cpu, _ = run_program(cpu, q)
        gpu.close(); cpu.close()
    a, b = res["3"], res["4"]
    assert a[1] == b[1] == st_cpu and a[3] == b[3]
    for x, y, z, nm in zip(a[0], b[0], o_cpu, names):
        assert np.array_equal(x, y, equal_nan=True), nm
        assert _rel(y, z) <= 1e-9, nm
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.parametrize("name,S,seed", [(n, 96, 11 + i) for i, n in enumerate(program_cases())] +
                         [("two_reservoir.route", 1024, 31), ("thames.route", 4096, 32)])
def test_specialised_run_kernel_equals_generic_run_kernel(monkeypatch, name, S, seed):
    """b200lp_load_program compiles a run kernel specialised on (structure, program) with NVRTC (csrc/jit_run.inc: ops, row
    bounds, Storage.after and recorders as straight-line code over registers, no shared memory on the steady-state timestep).
    Same arithmetic as the generic lp_run_kernel (B200LP_JIT=0): bit-identical recorders, final state and pivot counts, in
    one launch and in time slices; both equal to the oracle's."""
    s, prog, _, names = load_program_case(name)
    q = widen_program(s, prog, S, seed)
    res = {}
    for jit in ("1", "0"):
        monkeypatch.setenv("B200LP_JIT", jit)
        cpu, gpu = _engines(s, S)
        res[jit] = run_program(gpu, q, chunks=[7, 100] if jit == "1" else None) + (gpu.counters()["pivots"], gpu.stats()["run_kernel_jit"])
        if jit == "0":
            o_cpu, st_cpu, _ = run_program(cpu, q)
        gpu.close(); cpu.close()
    a, b = res["1"], res["0"]
    assert a[4] == 1 and b[4] == 0, "the specialised kernel must be the one that ran by default"
    assert a[1] == b[1] == st_cpu and a[3] == b[3]
    for x, y, z, nm in zip(a[0], b[0], o_cpu, names):
        assert np.array_equal(x, y, equal_nan=True), (nm, _rel(x, y))
        assert _rel(y, z) <= 1e-9, nm
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.parametrize("name,S,seed", [("two_reservoir.route", 301, 41), ("thames.route", 1000, 42), ("rolling_virtual_storage_long.route", 17, 43),
                                         ("lagged_flows.route", 45, 44)])
def test_sharded_cuda_engines_equal_one_engine(name, S, seed):
    """One host process, several engines (pywr_b200.engine.ShardedEngine): contiguous global_id blocks, one per GPU of the
    box — on a one-GPU box three handles on the same device.  Recorder arrays land in their column blocks through
    b200lp_fetch_recorder_strided; everything is bit-identical to the single engine."""
    from pywr_b200.engine import CudaEngine, ShardedEngine, load_library

    s, prog, _, names = load_program_case(name)
    q = widen_program(s, prog, S, seed)
    ndev = load_library().b200lp_device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    one = CudaEngine(s, S, device=0)
    many = ShardedEngine(s, S, devices=devices)
    o1, st1, x1 = run_program(one, q)
    o2, st2, x2 = run_program(many, q, chunks=[9, 77])
    assert st1 == st2
    for a, b, nm in zip(o1, o2, names):
        assert np.array_equal(a, b, equal_nan=True), nm
    for a, b in zip(x1, x2):
        assert np.array_equal(a, b, equal_nan=True)
    for r in range(q.n_recorders):
        if q.rec_shape(r, S) != (S,):
            for a, b in zip(one.fetch_recorder_agg(r), many.fetch_recorder_agg(r)):
                assert np.array_equal(a, b, equal_nan=True)
    assert many.counters()["pivots"] == one.counters()["pivots"]
    one.close(); many.close()
