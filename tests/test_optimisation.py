"""a whole population of candidate solutions evaluated in ONE device-resident batch
(pywr_b200.optimisation.PopulationEvaluator) returns, for every candidate, exactly what the reference's
``evaluate`` returns for it (pywr/optimisation/platypus.py:115-154: set the variable parameters, Model.run(),
aggregated_value() of the objective / constraint recorders).  CPU: oracle engine; GPU variant in
tests/test_gpu_pywr_dropin.py."""
import numpy as np
import pytest

pywr = pytest.importorskip("pywr")


def make_model_factory(n_inflow=3, years=1, solver="oracle"):
    from benchmarks.workloads import SEEDS, daily_index, synthetic_inflow, two_reservoir_document

    def make():
        from pywr.core import Scenario
        from pywr.model import Model
        from pywr.parameters import ArrayIndexedScenarioParameter

        idx = daily_index("2100-01-01", years=years)
        doc = two_reservoir_document(str(idx[0].date()), str(idx[-1].date()))
        doc["parameters"]["control_curve"].update(is_variable=True, lower_bounds=0.0, upper_bounds=1.0)
        doc["parameters"]["demand1_max"] = {"type": "constant", "value": 1.8, "is_variable": True,
                                            "lower_bounds": 0.5, "upper_bounds": 3.0}
        next(n for n in doc["nodes"] if n["name"] == "demand1")["max_flow"] = "demand1_max"
        doc["recorders"] = {
            "deficit1": {"type": "DeficitFrequencyNode", "node": "demand1", "is_objective": "minimise"},
            "transferred": {"type": "totalflownode", "node": "transfer", "is_objective": "maximise", "agg_func": "max"},
            "minvol2": {"type": "MinimumVolumeStorageRecorder", "node": "reservoir2", "constraint_lower_bounds": 10.0,
                        "agg_func": "min"},
            "volume1": {"type": "NumpyArrayStorageRecorder", "node": "reservoir1"},
        }
        m = Model.load(doc, solver=solver)
        axis = Scenario(m, "inflow", size=n_inflow)
        inflow = synthetic_inflow(idx.month.values, n_inflow, SEEDS["two_reservoir"])
        flow = ArrayIndexedScenarioParameter(m, axis, inflow, name="inflow")
        for c in ("catchment1", "catchment2"):
            m.nodes[c].flow = flow
        return m
    return make


def reference_evaluate(make, x):
    """One candidate the reference's way."""
    from pywr.optimisation import cache_constraints, cache_objectives, cache_variable_parameters

    m = make()
    m.setup()
    variables, vmap = cache_variable_parameters(m)
    for i, var in enumerate(variables):
        var.set_double_variables(np.ascontiguousarray(x[vmap[i]: vmap[i + 1]]))
    m.run()
    objs = [(1.0 if r.is_objective == "minimise" else -1.0) * r.aggregated_value() for r in cache_objectives(m)]
    cons = [c.aggregated_value() for c in cache_constraints(m)]
    return np.array(objs), np.array(cons)


def random_population(ev, P, seed):
    rng = np.random.default_rng(seed)
    X = np.empty((P, ev.n_variables))
    for i, var in enumerate(ev.variables):
        lo, hi = np.asarray(var.get_double_lower_bounds()), np.asarray(var.get_double_upper_bounds())
        X[:, ev.variable_map[i]: ev.variable_map[i + 1]] = rng.uniform(lo, hi, size=(P, len(lo)))
    return X


def check_population(engine_factory, P=5, n_inflow=3):
    import oracle.pywr_oracle_solver  # noqa: F401
    from pywr_b200.optimisation import PopulationEvaluator

    make = make_model_factory(n_inflow=n_inflow)
    ev = PopulationEvaluator(make, P, engine_factory=engine_factory)
    assert ev.n_variables == 13 and len(ev.objectives) == 2 and len(ev.constraints) == 1
    assert ev.S == P * n_inflow
    for gen in range(2):                                  # two generations through the same engine
        X = random_population(ev, P, seed=gen)
        objs, cons = ev.evaluate(X)
        assert objs.shape == (P, 2) and cons.shape == (P, 1)
        for p in range(P):
            o_ref, c_ref = reference_evaluate(make, X[p])
            assert np.allclose(objs[p], o_ref, rtol=1e-9, atol=1e-12), (gen, p, objs[p], o_ref)
            assert np.allclose(cons[p], c_ref, rtol=1e-9, atol=1e-12), (gen, p, cons[p], c_ref)
        assert np.ptp(objs[:, 0]) > 0 or np.ptp(objs[:, 1]) > 0   # candidates really differ
    ev.close()


def test_population_evaluator_matches_reference_evaluate_on_the_oracle():
    from oracle.oracle_engine import OracleEngine

    check_population(OracleEngine)
