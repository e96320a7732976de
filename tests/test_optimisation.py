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


class _FakePlatypus:
    """The slice of the Platypus API the reference's wrapper and ours use (platypus-opt is not installable here): Problem with
    types / constraints lists and a function, Real with bounds and decode, Constraint as a callable distance to feasibility,
    Solution, Evaluator, and an EvaluateJob-like object carrying a solution."""

    class Real:
        def __init__(self, lo, hi):
            self.min_value, self.max_value = float(lo), float(hi)

        def decode(self, v):
            return v

    class Constraint:
        def __init__(self, op, value=0.0):
            self.op, self.value = op, float(value)

        def __call__(self, x):
            if self.op == ">=":
                return 0.0 if x >= self.value else self.value - x
            if self.op == "<=":
                return 0.0 if x <= self.value else x - self.value
            return abs(x - self.value)

    class Problem:
        def __init__(self, nvars, nobjs, nconstrs=0):
            self.nvars, self.nobjs, self.nconstrs = nvars, nobjs, nconstrs
            self.types, self.constraints, self.function = [None] * nvars, [None] * nconstrs, None

    class Solution:
        def __init__(self, problem):
            self.problem = problem
            self.variables = [None] * problem.nvars
            self.objectives, self.constraints = [None] * problem.nobjs, [None] * problem.nconstrs
            self.constraint_violation, self.feasible, self.evaluated = 0.0, True, False

    class Evaluator:
        def evaluate_all(self, jobs, **kwargs):
            raise NotImplementedError

    class Job:
        def __init__(self, solution):
            self.solution = solution


def test_platypus_wrapper_evaluates_a_generation_in_batches():
    """PlatypusPopulationWrapper: the reference's platypus.Problem (pywr/optimisation/platypus.py:31-112), evaluated
    population_size candidates per device run through a platypus.Evaluator; every solution receives what the reference's
    one-Model.run()-per-candidate evaluate (:115-154) returns, plus platypus' constraint violation / feasibility."""
    import oracle.pywr_oracle_solver  # noqa: F401
    from oracle.oracle_engine import OracleEngine
    from pywr_b200.optimisation import PlatypusPopulationWrapper

    make = make_model_factory(n_inflow=2)
    P = 4
    w = PlatypusPopulationWrapper(make, P, platypus=_FakePlatypus, engine_factory=OracleEngine)
    pb = w.problem
    assert (pb.nvars, pb.nobjs, pb.nconstrs) == (13, 2, 1)
    assert all(isinstance(t, _FakePlatypus.Real) for t in pb.types) and sorted({t.max_value for t in pb.types}) == [1.0, 3.0]
    assert pb.constraints[0].op == ">=" and pb.constraints[0].value == 10.0
    X = random_population(w.batch, 6, seed=5)            # 6 solutions: one full batch of 4 and a padded one
    sols = []
    for x in X:
        s = _FakePlatypus.Solution(pb)
        s.variables[:] = list(x)
        sols.append(s)
    jobs = w.evaluator.evaluate_all([_FakePlatypus.Job(s) for s in sols])
    assert len(jobs) == 6 and w.batch.evaluations == 2
    for s, x in zip(sols, X):
        o_ref, c_ref = reference_evaluate(make, x)
        assert s.evaluated and np.allclose(s.objectives, o_ref, rtol=1e-9, atol=1e-12)
        assert np.allclose(s.constraints, c_ref, rtol=1e-9, atol=1e-12)
        assert s.constraint_violation == pytest.approx(max(0.0, 10.0 - c_ref[0]), abs=1e-9)
        assert s.feasible == (s.constraint_violation == 0.0)
    one = pb.function(list(X[2]))                           # single-solution entry point
    assert np.allclose(one[0], sols[2].objectives, rtol=1e-12) and np.allclose(one[1], sols[2].constraints, rtol=1e-12)
    w.close()
