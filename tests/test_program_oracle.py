"""CPU tests of the device-resident run semantics: the oracle's program mode against recorder arrays
produced by a HOST run of the reference framework (Model.run with the reference's own Parameter /
Storage.after / Recorder code; tests/golden/make_golden_programs.py)."""
import numpy as np
import pytest

from fixtures_util import rel_diff, load_program_case, n_scenarios_of, program_cases, run_program, widen_program
from oracle.oracle_engine import OracleEngine


def _rel(a, b):
    return rel_diff(a, b)


@pytest.mark.parametrize("name", program_cases())
def test_oracle_program_matches_reference_host_run(name):
    s, prog, expected, names = load_program_case(name)
    eng = OracleEngine(s, n_scenarios_of(s, prog))
    outs, status, _ = run_program(eng, prog)
    assert status[0] == 0
    T = prog.n_steps
    for r, (out, exp, nm) in enumerate(zip(outs, expected, names)):
        if nm.startswith(("mf_", "df_")):       # the reference divides by T in finish()
            out = out / T
        assert out.shape == exp.shape
        # schedule-driven outputs bit-exact in shape; values: same arithmetic, summation order of the
        # storage net flow differs in the last bit -> 1e-12
        assert _rel(out, exp) <= 1e-12, (nm, _rel(out, exp))
    eng.close()


def test_chunked_run_is_identical():
    s, prog, _, _ = load_program_case("thames.route")
    S = n_scenarios_of(s, prog)
    a = OracleEngine(s, S); b = OracleEngine(s, S)
    oa, _, sa = run_program(a, prog)
    ob, _, sb = run_program(b, prog, chunks=[1, 7, 100, 300])
    for x, y in zip(oa, ob):
        assert np.array_equal(x, y)
    for x, y in zip(sa, sb):
        assert np.array_equal(x, y)
    a.close(); b.close()


def test_widened_program_threads_identical():
    s, prog, _, _ = load_program_case("two_reservoir.route")
    q = widen_program(s, prog, 48, seed=5)
    a = OracleEngine(s, 48, threads=1); b = OracleEngine(s, 48, threads=6)
    oa, sta, _ = run_program(a, q)
    ob, stb, _ = run_program(b, q)
    assert sta == stb and sta[0] == 0
    for x, y in zip(oa, ob):
        assert np.array_equal(x, y)
    # scenarios really differ
    assert np.ptp(oa[0][-1]) > 0
    a.close(); b.close()


def test_run_on_device_with_pywr_objects():
    """Full host path: compile from live Pywr objects, run (oracle engine), write back through the
    Cython bridge, read the reference's recorders."""
    pywr = pytest.importorskip("pywr")
    import os
    ref = os.environ.get("PYWR_REFERENCE", "/root/reference")
    path = os.path.join(ref, "examples", "thames", "thames.json")
    if not os.path.exists(path):
        pytest.skip("reference model documents not available")
    import oracle.pywr_oracle_solver  # noqa: F401
    from pywr.model import Model
    from pywr.recorders import MeanFlowNodeRecorder, NumpyArrayNodeRecorder, NumpyArrayStorageRecorder
    from pywr_b200.device_run import run_on_device

    def make():
        m = Model.load(path, solver="oracle")
        m.timestepper.end = "2100-12-31"
        return m, [NumpyArrayStorageRecorder(m, m.nodes["reservoir1"]), NumpyArrayNodeRecorder(m, m.nodes["demand1"]),
                   MeanFlowNodeRecorder(m, m.nodes["demand1"])]
    m1, r1 = make(); m1.run()
    m2, r2 = make(); res = run_on_device(m2, engine_factory=OracleEngine)
    assert res.timesteps == 365 and res.num_scenarios == 5
    assert _rel(np.array(r2[0].data), np.array(r1[0].data)) <= 1e-12
    assert _rel(np.array(r2[1].data), np.array(r1[1].data)) <= 1e-12
    assert _rel(np.array(r2[2].values()), np.array(r1[2].values())) <= 1e-12
    assert np.allclose(m2.nodes["reservoir1"].volume, m1.nodes["reservoir1"].volume, rtol=1e-12)
    assert np.allclose(m2.nodes["demand1"].flow, m1.nodes["demand1"].flow, rtol=1e-12)
    assert r2[0].to_dataframe().shape == (365, 5)


def test_host_mirror_protocol_keeps_host_recorders_alive():
    """recorders the device cannot evaluate (a custom Python recorder reading node.flow /
    storage.volume every timestep) run on the host against mirrored state and see exactly what they see in a
    host run; device recorders of the same model are unaffected.  Parameter recorders (value and index) are
    device recorders: samples of the parameter's slot."""
    pytest.importorskip("pywr")
    import os
    ref = os.environ.get("PYWR_REFERENCE", "/root/reference")
    path = os.path.join(ref, "examples", "thames", "thames.json")
    if not os.path.exists(path):
        pytest.skip("reference model documents not available")
    import oracle.pywr_oracle_solver  # noqa: F401
    from pywr.model import Model
    from pywr.recorders import (NumpyArrayIndexParameterRecorder, NumpyArrayNodeRecorder, NumpyArrayParameterRecorder,
                                NumpyArrayStorageRecorder, Recorder)
    from pywr_b200.device_run import run_on_device
    from pywr_b200.program import Unsupported

    class FlowAndVolume(Recorder):
        def __init__(self, model, node, storage, **kw):
            super().__init__(model, **kw)
            self.node, self.storage = node, storage

        def setup(self):
            self.data = np.zeros((len(self.model.timestepper), len(self.model.scenarios.combinations)))

        def reset(self):
            self.data[...] = 0.0

        def after(self):
            ts = self.model.timestepper.current
            self.data[ts.index] = np.asarray(self.node.flow) ** 2 + np.asarray(self.storage.volume)

    def make():
        m = Model.load(path, solver="oracle")
        m.timestepper.end = "2100-06-30"
        level = [p for p in m.parameters if type(p).__name__ == "ControlCurveIndexParameter"][0]
        return m, [NumpyArrayStorageRecorder(m, m.nodes["reservoir1"]), NumpyArrayNodeRecorder(m, m.nodes["demand1"]),
                   FlowAndVolume(m, m.nodes["demand1"], m.nodes["reservoir1"], name="custom"),
                   NumpyArrayParameterRecorder(m, m.parameters["demand_max_flow"], name="param"),
                   NumpyArrayIndexParameterRecorder(m, level, name="level")], level

    m1, r1, _ = make(); m1.run()
    m2, r2, _ = make(); res = run_on_device(m2, engine_factory=OracleEngine)
    assert sorted(res.host_recorders) == ["custom"]
    for a, b in zip(r1, r2):
        assert _rel(np.array(b.data), np.array(a.data)) <= 1e-12, b.name
    assert np.array(r2[4].data).dtype == np.array(r1[4].data).dtype and np.array_equal(np.array(r2[4].data), np.array(r1[4].data))
    assert np.allclose(np.asarray(r2[3].values()), np.asarray(r1[3].values()), rtol=1e-12)
    assert np.ptp(np.array(r1[2].data)) > 0 and np.ptp(np.array(r1[3].data)) > 0 and np.ptp(np.array(r1[4].data)) > 0
    m3, _, _ = make()
    with pytest.raises(Unsupported):
        run_on_device(m3, engine_factory=OracleEngine, host_mirror=False)


@pytest.mark.parametrize("arrays", ["lazy", "none"])
def test_run_on_device_without_shipping_arrays(arrays):
    """arrays="lazy" / "none": the [T, S] arrays stay on (or never exist on) the device; Recorder.values() and
    aggregated_value() come from the temporal aggregates accumulated during the run and equal the host run's
    (pywr/recorders/_recorders.pyx:109-184,630-633); "lazy" still hands out any array on request."""
    pytest.importorskip("pywr")
    import os
    ref = os.environ.get("PYWR_REFERENCE", "/root/reference")
    path = os.path.join(ref, "examples", "thames", "thames.json")
    if not os.path.exists(path):
        pytest.skip("reference model documents not available")
    import oracle.pywr_oracle_solver  # noqa: F401
    from pywr.model import Model
    from pywr.recorders import MeanFlowNodeRecorder, NumpyArrayNodeRecorder, NumpyArrayStorageRecorder
    from pywr_b200.device_run import run_on_device
    from pywr_b200.engine import ShardedEngine
    from pywr_b200.program import Unsupported

    def make(extra=None):
        m = Model.load(path, solver="oracle")
        m.timestepper.end = "2100-12-31"
        recs = [NumpyArrayStorageRecorder(m, m.nodes["reservoir1"], temporal_agg_func="min", name="vmin"),
                NumpyArrayNodeRecorder(m, m.nodes["demand1"], temporal_agg_func="mean", agg_func="max", name="supply"),
                NumpyArrayNodeRecorder(m, m.nodes["demand1"], temporal_agg_func="sum", name="total"),
                NumpyArrayStorageRecorder(m, m.nodes["reservoir1"], temporal_agg_func="max", name="vmax"),
                MeanFlowNodeRecorder(m, m.nodes["demand1"], name="mf")]
        if extra:
            recs.append(NumpyArrayNodeRecorder(m, m.nodes["demand1"], temporal_agg_func=extra, name="extra"))
        return m, recs
    m1, r1 = make(); m1.run()
    # ... and through the one-process multi-device layer (two shards of the 5 scenarios)
    factory = lambda s, S: ShardedEngine(s, S, devices=2, factory=lambda st, n, d: OracleEngine(st, n, threads=1))  # noqa: E731
    for fac in (OracleEngine, factory):
        m2, r2 = make(); res = run_on_device(m2, engine_factory=fac, arrays=arrays)
        for a, b in zip(r1, r2):
            assert _rel(np.array(b.values()), np.array(a.values())) <= 1e-12, a.name
            assert abs(b.aggregated_value() - a.aggregated_value()) <= 1e-12 * max(1.0, abs(a.aggregated_value())), a.name
        assert np.asarray(r2[0].data).shape == (1, 5)          # one row: the temporal aggregate
        if arrays == "lazy":
            assert _rel(res.recorders.data("supply"), np.array(r1[1].data)) <= 1e-12
            assert np.array_equal(res.recorders.data("mf"), res.recorders["mf"])
            res.close()
            with pytest.raises(RuntimeError):
                res.recorders.data("vmin")
    # a temporal aggregation the device does not accumulate: fetched eagerly ("lazy") / refused ("none")
    m3, r3 = make("median")
    if arrays == "none":
        with pytest.raises(Unsupported):
            run_on_device(m3, engine_factory=OracleEngine, arrays=arrays)
    else:
        m4, r4 = make("median"); m4.run()
        res = run_on_device(m3, engine_factory=OracleEngine, arrays=arrays)
        assert np.asarray(r3[-1].data).shape == (365, 5)
        assert _rel(np.array(r3[-1].values()), np.array(r4[-1].values())) <= 1e-12
        res.close()


def test_constant_factor_changed_between_runs_reaches_the_lp():
    """Aggregated-node factors that are ConstantParameters are baked into the matrix at setup; the reference re-reads
    get_factors_norm every scenario and timestep (cython_glpk.pyx:663-673), so a factor given a new value between two runs
    (an optimisation variable) changes the flow ratio without the model being marked dirty.  The plug-in notices the change
    when it gathers the timestep's inputs and rebuilds."""
    pytest.importorskip("pywr")
    import oracle.pywr_oracle_solver  # noqa: F401
    from pywr.model import Model
    from pywr.nodes import AggregatedNode, Input, Output

    def make(f1):
        m = Model(solver="oracle")
        A = Input(m, "A", max_flow=10.0)
        B = Input(m, "B", max_flow=10.0)
        Z = Output(m, "Z", cost=-10)
        A.connect(Z); B.connect(Z)
        agg = AggregatedNode(m, "agg", [A, B])
        agg.max_flow = 9.0
        agg.factors = [1.0, f1]
        return m, A, B, agg
    m, A, B, agg = make(2.0)
    m.step()
    assert np.allclose([A.flow[0], B.flow[0]], [3.0, 6.0])
    assert not m.dirty
    agg.factors[1].set_double_variables(np.array([0.5]))     # the optimisation wrappers' way of changing a ConstantParameter
    assert not m.dirty
    m.step()
    m2, A2, B2, _ = make(0.5)
    m2.step()
    assert np.allclose([A.flow[0], B.flow[0]], [A2.flow[0], B2.flow[0]]) and np.allclose(B.flow[0], 3.0)


def test_model_run_with_device_run_solver_option():
    """The literal drop-in: a model document whose solver entry says ``device_run`` runs ENTIRELY on the device when the
    user calls ``Model.run()`` (pywr/_model.pyx:637-682) — same ModelResult, same recorder contents — and falls back to the
    per-timestep loop for a model the device cannot evaluate (a custom Python parameter here), unless "require" is given."""
    pytest.importorskip("pywr")
    import os
    ref = os.environ.get("PYWR_REFERENCE", "/root/reference")
    path = os.path.join(ref, "examples", "thames", "thames.json")
    if not os.path.exists(path):
        pytest.skip("reference model documents not available")
    import oracle.pywr_oracle_solver  # noqa: F401
    from pywr.model import Model
    from pywr.parameters import Parameter
    from pywr.recorders import NumpyArrayNodeRecorder, NumpyArrayStorageRecorder, TotalDeficitNodeRecorder
    from pywr_b200.program import Unsupported

    def make(**solver_args):
        m = Model.load(path, solver="oracle", solver_args=solver_args)
        m.timestepper.end = "2100-12-31"
        return m, [NumpyArrayStorageRecorder(m, m.nodes["reservoir1"]), NumpyArrayNodeRecorder(m, m.nodes["demand1"]),
                   TotalDeficitNodeRecorder(m, m.nodes["demand1"])]
    m1, r1 = make(); res1 = m1.run()
    m2, r2 = make(device_run=True)
    calls = []
    orig_solve = m2.solver.solve
    m2.solver.solve = lambda model: (calls.append(1), orig_solve(model))[1]
    res2 = m2.run()
    assert not calls, "no per-timestep solve may happen on a device run"
    assert res2.timesteps == res1.timesteps == 365 and res2.num_scenarios == 5 and res2.solver_settings["device_run"] is True
    assert res2.timestep.index == res1.timestep.index and "device_run_s" in res2.solver_stats
    for a, b in zip(r1, r2):
        assert _rel(np.asarray(b.values()), np.asarray(a.values())) <= 1e-12
    assert _rel(np.array(r2[0].data), np.array(r1[0].data)) <= 1e-12
    res2.to_dataframe()

    class Custom(Parameter):          # something only the host can evaluate
        def value(self, ts, si):
            return 3.5

    m3, r3 = make(device_run=True)
    m3.nodes["abs1"].max_flow = Custom(m3)
    res3 = m3.run()                   # falls back to the reference's loop
    assert res3.timesteps == 365 and "device_run" not in res3.solver_settings
    assert _rel(np.array(r3[0].data), np.array(r1[0].data)) <= 1e-12
    m4, _ = make(device_run="require")
    m4.nodes["abs1"].max_flow = Custom(m4)
    with pytest.raises(Unsupported):
        m4.run()


@pytest.mark.parametrize("arrays", ["eager", "none"])
def test_lagged_flow_parameters_without_a_user_recorder_on_the_node(arrays):
    """FlowParameter / NodeThresholdParameter read yesterday's flow of a node (op B200LP_OP_LAG).  The state is a flow recorder's
    [T, S] array: with no user recorder on the node the compiler appends a hidden one, and a read-out without arrays keeps
    storing them (no B200LP_PROG_AGG_ONLY)."""
    pytest.importorskip("pywr")
    import json
    import os
    ref = os.environ.get("PYWR_REFERENCE", "/root/reference")
    path = os.path.join(ref, "examples", "thames", "thames.json")
    if not os.path.exists(path):
        pytest.skip("reference model documents not available")
    import oracle.pywr_oracle_solver  # noqa: F401
    from pywr.model import Model
    from pywr.recorders import NumpyArrayStorageRecorder, TotalDeficitNodeRecorder
    from pywr_b200.device_run import run_on_device
    from pywr_b200.program import PROG_AGG_ONLY

    def make():
        doc = json.load(open(path))
        for prm in doc["parameters"].values():
            if isinstance(prm, dict) and "url" in prm and not os.path.isabs(prm["url"]):
                prm["url"] = os.path.join(ref, "examples/thames", prm["url"])
        P = doc["parameters"]
        P["river_yesterday"] = {"type": "flow", "node": "term1", "initial_value": 1.0}
        P["abs_cap"] = {"type": "aggregated", "agg_func": "min",
                        "parameters": [3.5, {"type": "aggregated", "agg_func": "sum", "parameters": ["river_yesterday", 0.5]}]}
        next(n for n in doc["nodes"] if n["name"] == "abs1")["max_flow"] = "abs_cap"
        P["short"] = {"type": "nodethreshold", "node": "demand1", "threshold": 1.6, "predicate": "LT", "values": [1.0, 0.93]}
        P["demand_max_flow"]["parameters"].append("short")
        m = Model.load(doc, solver="oracle")
        m.timestepper.end = "2100-12-31"
        return m, [NumpyArrayStorageRecorder(m, m.nodes["reservoir1"], name="vol"), TotalDeficitNodeRecorder(m, m.nodes["demand1"], name="tdef")]

    m1, r1 = make(); m1.run()
    m2, r2 = make(); res = run_on_device(m2, engine_factory=OracleEngine, arrays=arrays)
    assert res.program.n_recorders == 4 and len(res.program.recorders) == 2          # two hidden flow recorders
    assert res.program.has_lag and not (res.program.flags & PROG_AGG_ONLY)
    for a, b in zip(r1, r2):
        assert np.allclose(np.asarray(b.values()), np.asarray(a.values()), rtol=1e-12, atol=0), b.name
    if arrays == "eager":
        assert _rel(np.array(r2[0].data), np.array(r1[0].data)) <= 1e-12
    assert np.ptp(np.array(r1[0].data)) > 0


def test_duration_curve_and_level_recorders_run_on_the_device():
    """Subclasses that only post-process the recorded array in finish() (flow / storage duration curves) ride on the device's array
    recorders; NumpyArrayLevelRecorder samples the storage's level parameter (an InterpolatedVolumeParameter here); Total / Mean
    parameter recorders accumulate a parameter's slot."""
    pytest.importorskip("pywr")
    import os
    ref = os.environ.get("PYWR_REFERENCE", "/root/reference")
    path = os.path.join(ref, "examples", "thames", "thames.json")
    if not os.path.exists(path):
        pytest.skip("reference model documents not available")
    import oracle.pywr_oracle_solver  # noqa: F401
    from pywr.model import Model
    from pywr.parameters import InterpolatedVolumeParameter
    from pywr.recorders import (FlowDurationCurveRecorder, MeanParameterRecorder, NumpyArrayLevelRecorder, StorageDurationCurveRecorder,
                                TotalParameterRecorder)
    from pywr_b200.device_run import run_on_device

    def make():
        m = Model.load(path, solver="oracle")
        m.timestepper.end = "2100-12-31"
        res = m.nodes["reservoir1"]
        res.level = InterpolatedVolumeParameter(m, res, [0.0, 60.0, 140.0, 200.0], [10.0, 14.0, 17.5, 19.0],
                                                interp_kwargs={"bounds_error": False, "fill_value": (10.0, 19.0)})
        pct = [5.0, 25.0, 50.0, 75.0, 95.0]
        return m, [FlowDurationCurveRecorder(m, m.nodes["term1"], pct, name="fdc"),
                   StorageDurationCurveRecorder(m, res, pct, name="sdc"), NumpyArrayLevelRecorder(m, res, name="level"),
                   TotalParameterRecorder(m, m.parameters["demand_max_flow"], factor=2.0, integrate=True, name="total"),
                   TotalParameterRecorder(m, m.parameters["demand_max_flow"], name="plain_total"),
                   MeanParameterRecorder(m, m.parameters["demand_saving_factor"], factor=0.5, name="mean")]

    m1, r1 = make(); m1.run()
    m2, r2 = make(); out = run_on_device(m2, engine_factory=OracleEngine, arrays="lazy")
    assert out.host_recorders == []
    assert _rel(np.array(r2[0].fdc), np.array(r1[0].fdc)) <= 1e-12 and np.ptp(np.array(r1[0].fdc)) > 0
    assert _rel(np.array(r2[1].sdc), np.array(r1[1].sdc)) <= 1e-12
    assert np.allclose(np.asarray(r2[2].values()), np.asarray(r1[2].values()), rtol=1e-12)
    for a, b in zip(r1, r2):
        assert np.allclose(np.asarray(b.values()), np.asarray(a.values()), rtol=1e-12, equal_nan=True), b.name


def test_ratcheting_thresholds_run_on_the_device():
    """AbstractThresholdParameter(ratchet=True): once the predicate has fired the index stays 1 until the model is reset
    (pywr/parameters/_thresholds.pyx:95-112).  On the device the index of the day before is the lagged sample of a hidden recorder
    of the parameter's own slot."""
    pytest.importorskip("pywr")
    import json
    import os
    ref = os.environ.get("PYWR_REFERENCE", "/root/reference")
    path = os.path.join(ref, "examples", "thames", "thames.json")
    if not os.path.exists(path):
        pytest.skip("reference model documents not available")
    import oracle.pywr_oracle_solver  # noqa: F401
    from pywr.model import Model
    from pywr.parameters import RecorderThresholdParameter
    from pywr.recorders import NumpyArrayIndexParameterRecorder, NumpyArrayNodeRecorder, NumpyArrayStorageRecorder
    from pywr_b200.device_run import run_on_device

    def make():
        doc = json.load(open(path))
        for prm in doc["parameters"].values():
            if isinstance(prm, dict) and "url" in prm and not os.path.isabs(prm["url"]):
                prm["url"] = os.path.join(ref, "examples/thames", prm["url"])
        P = doc["parameters"]
        mv = next(n for n in doc["nodes"] if n["name"] == "reservoir1")["max_volume"]
        P["drought_order"] = {"type": "storagethreshold", "storage_node": "reservoir1", "threshold": 0.7 * mv, "predicate": "LT",
                              "values": [1.0, 0.8], "ratchet": True}
        P["river_low"] = {"type": "nodethreshold", "node": "term1", "threshold": 1.0, "predicate": "LT", "values": [1.0, 0.97], "ratchet": True}
        P["demand_max_flow"]["parameters"] += ["drought_order", "river_low"]
        m = Model.load(doc, solver="oracle")
        m.timestepper.end = "2101-12-31"
        recs = [NumpyArrayStorageRecorder(m, m.nodes["reservoir1"], name="vol"), NumpyArrayNodeRecorder(m, m.nodes["demand1"], name="supplied"),
                NumpyArrayIndexParameterRecorder(m, m.parameters["drought_order"], name="order"),
                NumpyArrayIndexParameterRecorder(m, m.parameters["river_low"], name="low")]
        # yesterday's row of an array recorder against a threshold (RecorderThresholdParameter, :320-360; initial_value on day one)
        by_rec = RecorderThresholdParameter(m, recs[0], 150.0, values=[1.0, 0.98], predicate="LT", initial_value=0, name="vol_was_low")
        m.parameters["demand_max_flow"].add(by_rec)
        return m, recs + [NumpyArrayIndexParameterRecorder(m, by_rec, name="by_rec")]

    m1, r1 = make(); m1.run()
    m2, r2 = make(); out = run_on_device(m2, engine_factory=OracleEngine)
    assert out.host_recorders == []
    for a, b in zip(r1, r2):
        assert _rel(np.array(b.data, dtype=float), np.array(a.data, dtype=float)) <= 1e-12, b.name
    assert np.ptp(np.array(r1[4].data)) > 0
    order = np.array(r1[2].data)
    assert order.min() == 0 and order.max() == 1 and np.all(np.diff(order, axis=0) >= 0)      # it did fire, and never released
