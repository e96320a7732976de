"""BASELINE.json configs[0]: examples/two_reservoir, ONE scenario, daily timestep over the model's
whole 100-year record (T = 36,523).  The fixture (tests/golden/make_config1.py) holds a host run of the reference framework
(its own Parameter / Storage.after / Recorder code) and the HiGHS optimum of the LP at 96 timesteps spread over the run."""
import os

import numpy as np
import pytest

from fixtures_util import GOLDEN, rel_diff, run_program
from pywr_b200.program import ARRAY_KINDS, Program
from pywr_b200.structure import LPStructure

BASE = os.path.join(GOLDEN, "config1", "route")


def _load():
    s = LPStructure.load(BASE + ".pstructure.npz")
    prog = Program.load(BASE + ".program.npz")
    z = np.load(BASE + ".expected.npz")
    return s, prog, z


def _check_against_host_run(prog, z, outs, tol):
    stride = int(z["stride"])
    assert prog.n_steps == 36523                     # schedule-driven output: bit-exact
    for i, out in enumerate(outs):
        name = str(z["names"][i])
        if int(prog.rec_kind[i]) in ARRAY_KINDS:
            assert out.shape == (prog.n_steps, 1)
            assert rel_diff(out[::stride], z[f"rec{i}"]) <= tol, name
            agg = np.stack([np.cumsum(out, axis=0)[-1], out.min(axis=0), out.max(axis=0)])
            assert rel_diff(agg, z[f"agg{i}"]) <= tol, name
        else:
            v = out / prog.n_steps if name.startswith(("mf_", "df_")) else out   # the reference divides in finish()
            assert rel_diff(v, z[f"rec{i}"]) <= tol, name


def test_oracle_runs_config1_like_the_reference_host_run():
    from oracle.oracle_engine import OracleEngine

    s, prog, z = _load()
    eng = OracleEngine(s, 1)
    outs, status, _ = run_program(eng, prog)
    assert status[0] == 0
    _check_against_host_run(prog, z, outs, 1e-12)
    eng.close()


def test_oracle_objective_equals_highs_on_config1():
    from oracle.oracle_engine import OracleEngine

    s = LPStructure.load(BASE + ".structure.npz")
    tr = np.load(BASE + ".trace.npz")
    n = tr["slots"].shape[0]
    eng = OracleEngine(s, n)
    slots = np.ascontiguousarray(tr["slots"][:, :, 0].T)           # every sampled timestep becomes one scenario of a batch
    assert np.all(tr["days"] == 1.0)
    obj = np.zeros((1, n))
    eng.reset(None)
    assert eng.step(1.0, slots, None, None, obj) == 0
    assert rel_diff(obj[0], tr["objective_highs"]) <= 1e-7           # north_star: objectives within 1e-7 relative
    eng.close()


@pytest.mark.gpu
def test_cuda_runs_config1_like_the_reference_host_run_and_the_oracle():
    """S = 1: one warp on the whole GPU, 36,523 dependent timesteps in one launch (specialised run kernel), and the same in
    the generic kernel; flows and volumes against the reference host run (north_star: 1e-6 relative; asserted 1e-9)."""
    from oracle.oracle_engine import OracleEngine
    from pywr_b200.engine import CudaEngine

    s, prog, z = _load()
    cpu, gpu = OracleEngine(s, 1), CudaEngine(s, 1, device=0)
    o_cpu, st_cpu, x_cpu = run_program(cpu, prog)
    o_gpu, st_gpu, x_gpu = run_program(gpu, prog)
    assert st_gpu == st_cpu and st_gpu[0] == 0
    assert gpu.stats()["run_kernel_jit"] == 1
    for a, b in zip(o_gpu, o_cpu):
        assert rel_diff(a, b) <= 1e-9
    assert gpu.counters()["pivots"] == cpu.counters()["pivots"]
    _check_against_host_run(prog, z, o_gpu, 1e-9)
    gpu.close(); cpu.close()


@pytest.mark.gpu
def test_cuda_objective_equals_highs_on_config1():
    from pywr_b200.engine import CudaEngine

    s = LPStructure.load(BASE + ".structure.npz")
    tr = np.load(BASE + ".trace.npz")
    n = tr["slots"].shape[0]
    gpu = CudaEngine(s, n, device=0)
    slots = gpu.alloc_planes(max(s.n_slots, 1))
    slots[: s.n_slots] = tr["slots"][:, :, 0].T
    obj = gpu.alloc_planes(1)
    gpu.reset(None)
    assert gpu.step(1.0, slots, None, None, obj) == 0
    assert rel_diff(obj[0], tr["objective_highs"]) <= 1e-7
    gpu.close()
