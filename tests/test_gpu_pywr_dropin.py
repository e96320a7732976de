"""GPU tests (-m gpu) of the drop-in boundary as a Pywr user sees it: the reference framework
(baseline/_ref) drives the CUDA engine through its own solver API — Model(solver="b200").run() —
and through the device-resident run; results are compared with the same model solved by the CPU
oracle plugged in the same way."""
import numpy as np
import pytest

from fixtures_util import rel_diff

pytestmark = pytest.mark.gpu
pywr = pytest.importorskip("pywr")


def _recorder_values(model):
    out = {}
    for rec in model.recorders:
        if type(rec).__name__ == "AggregatedRecorder":
            out[rec.name] = np.array([rec.aggregated_value()])
        elif hasattr(rec, "data"):
            out[rec.name] = np.array(rec.data)
        else:
            out[rec.name] = np.array(rec.values())
    return out


@pytest.mark.parametrize("name,S", [("thames", 7), ("two_reservoir", 33)])
def test_model_run_with_b200_solver_matches_oracle_solver(name, S):
    import oracle.pywr_oracle_solver  # noqa: F401
    import pywr_b200  # noqa: F401  (registers "b200")
    from benchmarks.workloads import build_model

    m_cpu = build_model(name, S, years=1, solver="oracle")
    m_gpu = build_model(name, S, years=1, solver="b200")
    assert m_gpu.solver.name == "b200"
    r_cpu = m_cpu.run()
    r_gpu = m_gpu.run()
    assert r_gpu.timestep.index == r_cpu.timestep.index == 364          # schedule: bit-exact
    assert r_gpu.num_scenarios == S and r_gpu.solver_name == "b200"
    assert r_gpu.solver_stats["number_of_rows"] == r_cpu.solver_stats["number_of_rows"] > 0
    a, b = _recorder_values(m_gpu), _recorder_values(m_cpu)
    assert a.keys() == b.keys()
    for k in a:
        assert rel_diff(a[k], b[k]) <= 1e-9, k
    for n_gpu, n_cpu in zip(m_gpu.nodes, m_cpu.nodes):
        assert rel_diff(n_gpu.flow, n_cpu.flow) <= 1e-9, n_gpu.name
    st = m_gpu.solver._engine.stats()
    assert st["kernel_launches"] >= 365 and st["h2d_bytes"] > 0


@pytest.mark.parametrize("name,S", [("thames", 64), ("two_reservoir", 100)])
def test_run_on_device_matches_host_run(name, S):
    import oracle.pywr_oracle_solver  # noqa: F401
    from benchmarks.workloads import build_model
    from pywr_b200.device_run import run_on_device

    m_host = build_model(name, S, years=1, solver="oracle")
    m_host.run()
    m_dev = build_model(name, S, years=1, solver="null")
    res = run_on_device(m_dev)
    assert res.timesteps == 365 and res.num_scenarios == S
    a, b = _recorder_values(m_dev), _recorder_values(m_host)
    for k in b:
        assert rel_diff(a[k], b[k]) <= 1e-9, k
    for n_dev, n_host in zip(m_dev.nodes, m_host.nodes):
        assert rel_diff(n_dev.flow, n_host.flow) <= 1e-9, n_dev.name
        if hasattr(n_host, "volume"):
            assert rel_diff(n_dev.volume, n_host.volume) <= 1e-9
    df = m_dev.recorders[next(iter(b))].to_dataframe() if hasattr(m_dev.recorders[next(iter(b))], "to_dataframe") else None
    assert df is None or df.shape[1] == S


def test_infeasible_model_raises_like_the_reference():
    import pywr_b200  # noqa: F401
    from pywr.core import Input, Model, Output

    model = Model(solver="b200")
    a = Input(model, "a", min_flow=10.0, max_flow=10.0)
    b = Output(model, "b", max_flow=5.0)
    a.connect(b)
    with pytest.raises(RuntimeError, match="Simplex solver failed"):
        model.run()


def test_nan_raises_lp_error():
    import pywr_b200  # noqa: F401
    from pywr.core import Input, Model, Output
    from pywr_b200.solver import B200LPError

    model = Model(solver="b200")
    a = Input(model, "a", max_flow=float("nan"))
    b = Output(model, "b", max_flow=5.0, cost=-1)
    a.connect(b)
    with pytest.raises(B200LPError):
        model.run()


def test_host_mirror_protocol_on_the_gpu():
    """SURVEY 8f rank 2 on the CUDA engine: a custom Python recorder stays on the host, fed by per-timestep mirrors of
    node flows / storage volumes; device recorders — the parameter recorder among them — run on the device."""
    import oracle.pywr_oracle_solver  # noqa: F401
    from benchmarks.workloads import build_model
    from pywr.recorders import NumpyArrayParameterRecorder, Recorder
    from pywr_b200.device_run import run_on_device

    class FlowAndVolume(Recorder):
        def __init__(self, model, node, storage, **kw):
            super().__init__(model, **kw)
            self.node, self.storage = node, storage

        def setup(self):
            self.data = np.zeros((len(self.model.timestepper), len(self.model.scenarios.combinations)))

        def reset(self):
            self.data[...] = 0.0

        def after(self):
            self.data[self.model.timestepper.current.index] = np.asarray(self.node.flow) + 0.5 * np.asarray(self.storage.volume)

    def make(solver):
        m = build_model("two_reservoir", 24, years=1, solver=solver)
        storage = [n for n in m.nodes if type(n).__name__ == "Storage"][0]
        node = [n for n in m.nodes if type(n).__name__ == "Output"][0]
        param = [p for p in m.parameters if type(p).__name__ == "ControlCurveParameter"][0]
        FlowAndVolume(m, node, storage, name="custom")
        NumpyArrayParameterRecorder(m, param, name="param")
        return m

    m_host = make("oracle"); m_host.run()
    m_dev = make("null"); res = run_on_device(m_dev)
    assert sorted(res.host_recorders) == ["custom"]      # the parameter recorder samples its slot on the device
    a, b = _recorder_values(m_dev), _recorder_values(m_host)
    for k in b:
        assert rel_diff(a[k], b[k]) <= 1e-9, k
    assert np.ptp(b["custom"]) > 0


def test_population_evaluator_on_the_gpu():
    """SURVEY 8f rank 4 on the CUDA engine: 12 candidates x 4 inflow scenarios in one batch, every candidate's
    objectives / constraint equal to the reference's per-candidate evaluate (host run with the oracle solver)."""
    from test_optimisation import check_population

    check_population(None, P=12, n_inflow=4)


def _network_model(solver, n_reaches=24, end="2100-03-31"):
    """generated river network (benchmarks.workloads.synthetic_network_document) with a 3-member wetness scenario and
    a few recorders; the same document as tests/golden/make_network_program.py at a size the oracle solves in seconds"""
    from pywr.model import Model
    from pywr.recorders import (MinimumVolumeStorageRecorder, NumpyArrayNodeRecorder, NumpyArrayStorageRecorder,
                                TotalDeficitNodeRecorder)

    from benchmarks.workloads import synthetic_network_document

    doc = synthetic_network_document(n_reaches=n_reaches, end=end)
    doc["scenarios"] = [{"name": "inflow", "size": 3}]
    params = doc["parameters"]
    params["wet"] = {"type": "constantscenario", "scenario": "inflow", "values": [1.0, 0.6, 1.3]}
    for name in [k for k in params if k.startswith("inflow")]:
        params[name] = {"type": "aggregated", "agg_func": "product", "parameters": [params[name], "wet"]}
    m = Model.load(doc, solver=solver)
    for n in m.nodes:
        cn = type(n).__name__
        if cn == "Storage":
            NumpyArrayStorageRecorder(m, n, name=f"vol_{n.name}")
            MinimumVolumeStorageRecorder(m, n, name=f"minv_{n.name}")
        elif n.name.startswith(("town", "farm")):
            TotalDeficitNodeRecorder(m, n, name=f"tdef_{n.name}")
    NumpyArrayNodeRecorder(m, m.nodes["sea"], name="flow_sea")
    return m


def test_large_lp_path_as_a_pywr_user_sees_it(monkeypatch):
    """The large-LP path (front end of pdhg_engine.cuh + revised simplex) behind both drop-in boundaries, on a generated
    network in the edge formulation: Model(solver="b200-edge").run() — Pywr's own before()/after() around one
    b200lp_step per timestep — and run_on_device(model) — the whole run as plane kernels around the batched solve —
    against the same model solved by the CPU oracle."""
    import oracle.pywr_oracle_solver  # noqa: F401
    import pywr_b200  # noqa: F401
    from pywr_b200.device_run import run_on_device

    monkeypatch.setenv("B200LP_FORCE_PDHG", "1")
    monkeypatch.setenv("B200LP_FORCE_RSX", "1")
    m_host = _network_model("oracle-edge")
    m_host.run()
    want = _recorder_values(m_host)
    m_step = _network_model("b200-edge")
    m_step.run()
    assert m_step.solver._engine.stats()["kernel_variant"] == 300
    got = _recorder_values(m_step)
    for k in want:
        assert rel_diff(got[k], want[k]) <= 1e-6, k            # north_star: flows / volumes within 1e-6 relative
    m_dev = _network_model("null")
    res = run_on_device(m_dev, formulation="edge")
    assert res.num_scenarios == 3 and res.timesteps == 90
    got = _recorder_values(m_dev)
    for k in want:
        assert rel_diff(got[k], want[k]) <= 1e-6, k
    for n_dev, n_host in zip(m_dev.nodes, m_host.nodes):
        if hasattr(n_host, "volume"):
            assert rel_diff(n_dev.volume, n_host.volume) <= 1e-6, n_dev.name


def test_mid_size_models_run_device_resident_without_any_override():
    """LP 158 x 104 (edge form of the 24-reach network): too large for the register-resident variants, small enough for the
    CTA-per-scenario simplex, which only serves b200lp_step.  run_on_device creates its engine with
    B200LP_CREATE_FOR_PROGRAM, so the model takes the large-LP path (revised simplex, variant 300) and still runs as a
    device-resident program; Model.run(solver="b200-edge") keeps the CTA engine (variant 100)."""
    import oracle.pywr_oracle_solver  # noqa: F401
    import pywr_b200  # noqa: F401
    from pywr_b200.device_run import run_on_device

    m_host = _network_model("oracle-edge")
    m_host.run()
    want = _recorder_values(m_host)
    m_dev = _network_model("null")
    res = run_on_device(m_dev, formulation="edge")
    assert res.counters["kernel_variant"] == 300
    got = _recorder_values(m_dev)
    for k in want:
        assert rel_diff(got[k], want[k]) <= 1e-6, k
    m_step = _network_model("b200-edge")
    m_step.run()
    assert m_step.solver._engine.stats()["kernel_variant"] == 100
    got = _recorder_values(m_step)
    for k in want:
        assert rel_diff(got[k], want[k]) <= 1e-6, k


def test_model_run_on_the_device_through_the_solver_option():
    """The literal drop-in on the GPU: ``Model.load(doc, solver="b200", solver_args={"device_run": True}).run()`` executes the
    whole run on the device (one launch of the specialised run kernel) and leaves Pywr's recorder objects as the per-timestep
    ``solver="b200"`` run leaves them; ``device_run_devices`` shards the scenarios over handles / GPUs."""
    from benchmarks.workloads import build_model

    S = 96
    host = build_model("thames", S, years=2, solver="b200")
    rh = host.run()
    dev = build_model("thames", S, years=2, solver="b200", solver_args={"device_run": "require"})
    rd = dev.run()
    assert rd.timesteps == rh.timesteps and rd.num_scenarios == S and rd.solver_settings.get("device_run") is True
    for name in ("volume", "supplied", "deficit"):
        a, b = host.recorders[name], dev.recorders[name]
        assert np.allclose(np.asarray(b.values()), np.asarray(a.values()), rtol=1e-9, atol=1e-12), name
    assert np.allclose(np.array(dev.recorders["volume"].data), np.array(host.recorders["volume"].data), rtol=1e-9, atol=1e-12)
    assert rd.solver_stats["device_run_s"] < rh.time_taken / 20      # the run itself; the first call also compiles the kernel
    many = build_model("thames", S, years=2, solver="b200", solver_args={"device_run": "require", "device_run_devices": [0, 0]})
    many.run()
    assert np.array_equal(np.array(many.recorders["volume"].data), np.array(dev.recorders["volume"].data))
