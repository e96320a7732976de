"""GPU parity tests (-m gpu): the CUDA engine through the C-ABI against the CPU oracle on the same
inputs.  Bars (BASELINE.json north_star): objectives 1e-7 relative, flows 1e-6 relative; because both
sides run the same pivoting rules with the same fma placement we assert 1e-9 and report the maximum."""
import numpy as np
import pytest

from fixtures_util import rel_diff, golden_cases, load_case, replay, synth_batch

pytestmark = pytest.mark.gpu


def _engines(s, S):
    from oracle.oracle_engine import OracleEngine
    from pywr_b200.engine import CudaEngine

    return OracleEngine(s, S, threads=8), CudaEngine(s, S, device=0)


def _rel(a, b):
    return rel_diff(a, b)


@pytest.mark.parametrize("name", golden_cases())
def test_golden_traces(name):
    """Replay of the traces recorded from the reference framework: CUDA == golden vectors."""
    s, tr = load_case(name)
    S = tr["slots"].shape[2]
    cpu, gpu = _engines(s, S)
    nf, cf, ob, nbad = replay(gpu, s, tr["slots"], tr["days"])
    assert nbad.sum() == 0
    assert _rel(ob, tr["objective"]) <= 1e-9
    assert _rel(nf, tr["node_flow"]) <= 1e-9
    assert _rel(cf, tr["col_flow"]) <= 1e-9
    assert _rel(ob, tr["objective_highs"]) <= 1e-7
    gpu.close(); cpu.close()


@pytest.mark.parametrize("name,S,T", [
    ("two_reservoir.route", 1024, 40), ("two_reservoir.edge", 512, 20), ("thames.route", 4096, 40),
    ("thames.edge", 1000, 20), ("three_storage.route", 777, 20), ("river_split_with_gauge1.route", 300, 20),
    ("loss_link_parameter.route", 513, 20), ("virtual_storage1.route", 100, 20), ("timeseries3.route", 2048, 30),
    ("reservoir2.route", 33, 10), ("piecewise1.edge", 7, 5),
])
def test_synthetic_batches(name, S, T):
    """Perturbed scenario batches (ragged sizes included): flows, objectives, basis and failure
    status agree between CUDA and oracle at every timestep."""
    s, tr = load_case(name)
    slots, days = synth_batch(s, tr, S, T, seed=7)
    cpu, gpu = _engines(s, S)
    r_cpu = replay(cpu, s, slots, days)
    r_gpu = replay(gpu, s, slots, days)
    assert np.array_equal(r_cpu[3] > 0, r_gpu[3] > 0)
    good = r_cpu[3] == 0
    assert _rel(r_gpu[2][good], r_cpu[2][good]) <= 1e-9
    assert _rel(r_gpu[0][good], r_cpu[0][good]) <= 1e-9
    assert _rel(r_gpu[1][good], r_cpu[1][good]) <= 1e-9
    if good.all():
        rb_c, cb_c = cpu.get_basis()
        rb_g, cb_g = gpu.get_basis()
        assert np.array_equal(rb_c, rb_g) and np.array_equal(cb_c, cb_g)
        assert cpu.counters()["pivots"] == gpu.counters()["pivots"]
    gpu.close(); cpu.close()


def test_status_codes_match_oracle():
    s, tr = load_case("reservoir2.route")
    cpu, gpu = _engines(s, 8)
    slots = np.repeat(tr["slots"][0], 8, axis=1)
    slots[0, 3] = np.nan            # NaN volume in scenario 3
    out_c = cpu.alloc_planes(s.n_nodes); out_g = gpu.alloc_planes(s.n_nodes)
    sl_g = gpu.alloc_planes(s.n_slots); sl_g[:] = slots
    cpu.reset(None); gpu.reset(None)
    assert cpu.step(1.0, slots, out_c, None, None) == 1
    assert gpu.step(1.0, sl_g, out_g, None, None) == 1
    assert cpu.status() == gpu.status() == (3, 4)
    ok = [k for k in range(8) if k != 3]
    assert np.array_equal(out_c[:, ok], out_g[:, ok])
    gpu.close(); cpu.close()


def test_reset_restarts_from_the_standard_basis():
    s, tr = load_case("thames.route")
    S = tr["slots"].shape[2]
    cpu, gpu = _engines(s, S)
    a = replay(gpu, s, tr["slots"][:20], tr["days"][:20])
    b = replay(gpu, s, tr["slots"][:20], tr["days"][:20])   # replay() resets first
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    gpu.close(); cpu.close()


@pytest.mark.parametrize("name,S,T", [("network24.edge", 96, 25), ("network12.route", 200, 25)])
def test_cta_kernel_on_larger_networks(name, S, T):
    """LPs beyond the register-resident variants run on the CTA-per-scenario shared-memory kernel."""
    s, tr = load_case(name)
    slots, days = synth_batch(s, tr, S, T, seed=13)
    cpu, gpu = _engines(s, S)
    assert gpu.stats()["kernel_variant"] == 100
    r_cpu = replay(cpu, s, slots, days)
    r_gpu = replay(gpu, s, slots, days)
    assert np.array_equal(r_cpu[3] > 0, r_gpu[3] > 0)
    good = r_cpu[3] == 0
    for k in range(3):
        assert _rel(r_gpu[k][good], r_cpu[k][good]) <= 1e-9
    if good.all():
        rb_c, cb_c = cpu.get_basis(); rb_g, cb_g = gpu.get_basis()
        assert np.array_equal(rb_c, rb_g) and np.array_equal(cb_c, cb_g)
        assert cpu.counters()["pivots"] == gpu.counters()["pivots"]
    gpu.close(); cpu.close()


@pytest.mark.parametrize("name", ["two_reservoir.route", "thames.edge", "loss_link_parameter.route", "virtual_storage1.route"])
def test_cta_kernel_forced_on_small_models(name, monkeypatch):
    """The same golden traces through the CTA kernel (B200LP_FORCE_CTA): both kernel families agree."""
    monkeypatch.setenv("B200LP_FORCE_CTA", "1")
    s, tr = load_case(name)
    S = tr["slots"].shape[2]
    cpu, gpu = _engines(s, S)
    assert gpu.stats()["kernel_variant"] == 100
    nf, cf, ob, nbad = replay(gpu, s, tr["slots"], tr["days"])
    assert nbad.sum() == 0
    assert _rel(nf, tr["node_flow"]) <= 1e-9 and _rel(cf, tr["col_flow"]) <= 1e-9 and _rel(ob, tr["objective"]) <= 1e-9
    gpu.close(); cpu.close()
