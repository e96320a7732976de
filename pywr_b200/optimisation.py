"""Population-at-once evaluation of a Pywr optimisation problem (BASELINE.json config 5).

The reference's optimisation wrappers evaluate ONE candidate solution per ``Model.run()``
(``pywr/optimisation/platypus.py:115-154``: set the variable parameters, run every scenario, read
``aggregated_value()`` of the objective / constraint recorders).  Here the whole population of a
generation becomes one more scenario axis: every variable parameter turns into a table with one column
per candidate (``program._Compiler``, ``population_axis``), all ``P x S`` combinations run in one
device-resident batch, and the objectives of candidate ``p`` are the recorder values aggregated over the
combinations that carry ``p`` — with the reference's own ``Aggregator`` so that the numbers are what
``evaluate`` returns for that candidate.

    ev = PopulationEvaluator(make_model, population_size=256)      # make_model() -> pywr Model
    objectives, constraints = ev.evaluate(X)                       # X: [P, n_variables]

Supported variable parameters: ``ConstantParameter`` and ``MonthlyProfileParameter`` (double variables).
"""

from __future__ import annotations

import numpy as np

from .program import ARRAY_KINDS, Unsupported, compile_program

POPULATION_AXIS = "b200_population"


class PopulationEvaluator:
    def __init__(self, make_model, population_size, engine_factory=None, formulation="route"):
        from pywr.core import Scenario
        from pywr.optimisation import cache_constraints, cache_objectives, cache_variable_parameters
        from .bridge import _bridge

        self._bridge = _bridge
        self.P = int(population_size)
        self.model = model = make_model()
        Scenario(model, POPULATION_AXIS, size=self.P)
        model.setup()
        self.variables, self.variable_map = cache_variable_parameters(model)
        self.objectives = cache_objectives(model)
        self.constraints = cache_constraints(model)
        for var in self.variables:
            if var.integer_size:
                raise Unsupported(f"integer variables ({var.name!r}) in a population batch")
        self.axis = [sc.name for sc in model.scenarios.scenarios].index(POPULATION_AXIS)
        self.structure, self.program = compile_program(model, formulation, population_axis=self.axis)
        missing = [v.name for v in self.variables if id(v) not in self.program.variable_tables]
        if missing:
            raise Unsupported(f"variable parameters {missing} do not feed the LP or a device recorder")
        # objectives need Recorder.values() only: when every array recorder aggregates over time with sum / mean / min /
        # max the device keeps those aggregates and the [T, S] arrays are neither stored nor shipped (B200LP_PROG_AGG_ONLY)
        from .device_run import temporal_aggregate
        from .program import PROG_AGG_ONLY

        dummy = np.zeros(1)
        self._agg_only = all(int(k) not in ARRAY_KINDS or temporal_aggregate(rec, 1, dummy, dummy, dummy) is not None
                             for k, rec in zip(self.program.rec_kind, self.program.recorders))
        if self._agg_only:
            if not self.program.has_lag:
                self.program.flags |= PROG_AGG_ONLY
        self.S = len(model.scenarios.combinations)
        self.member = np.array([np.asarray(si.indices)[self.axis] for si in model.scenarios.combinations])
        if engine_factory is None:
            from .engine import CudaEngine

            def engine_factory(structure, n_scenarios):
                return CudaEngine(structure, n_scenarios, for_program=True)
        self.engine = engine_factory(self.structure, self.S)
        self.engine.load_program(self.program)
        self.n_variables = self.variable_map[-1]
        self.evaluations = 0

    # -- one generation -----------------------------------------------------------------------------
    def evaluate(self, X):
        """``X[p]`` = the flat variable vector of candidate ``p`` in the reference's order
        (``cache_variable_parameters``).  Returns ``(objectives [P, n_obj], constraints [P, n_con])`` with the
        reference's sign convention (maximised objectives negated) and duplication of double-bounded
        constraints."""
        from pywr.recorders._recorders import Aggregator

        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.shape != (self.P, self.n_variables):
            raise ValueError(f"population must have shape {(self.P, self.n_variables)}, got {X.shape}")
        prog, eng = self.program, self.engine
        T = prog.n_steps
        for i, var in enumerate(self.variables):
            cols = X[:, self.variable_map[i]: self.variable_map[i + 1]]
            k = prog.variable_tables[id(var)]
            if type(var).__name__ == "ConstantParameter":
                data = cols[:, 0].reshape(1, self.P)
            else:                                   # MonthlyProfileParameter: value(ts) = values[ts.month - 1]
                data = cols[:, prog.months - 1].T   # [T, P]
            assert data.shape == (int(prog.table_rows[k]), int(prog.table_cols[k])), (data.shape, k)
            eng.update_table(k, np.ascontiguousarray(data))
        eng.program_reset()
        eng.run(0, T)
        eng.sync()
        n_failed, scen, code, tstep = eng.program_status()
        if n_failed:
            raise RuntimeError(f"{n_failed} scenario(s) failed; first: scenario {scen} (candidate {self.member[scen]}), "
                               f"status {code}, timestep {tstep}")
        from .device_run import temporal_aggregate

        for r, rec in enumerate(prog.recorders):
            if int(prog.rec_kind[r]) in ARRAY_KINDS and self._agg_only:
                row = temporal_aggregate(rec, T, *eng.fetch_recorder_agg(r))
                self._bridge.replace_array_recorder(rec, np.ascontiguousarray(row[None, :]))
                continue
            out = eng.fetch_recorder(r, prog.rec_shape(r, self.S))
            if int(prog.rec_kind[r]) in ARRAY_KINDS:
                self._bridge.set_array_recorder(rec, out)
            else:
                self._bridge.set_constant_recorder(rec, out)
        # MeanFlow / DeficitFrequency recorders divide by timestepper.current.index in finish() (parked past the end)
        from .device_run import park_timestepper

        park_timestepper(self.model)
        for rec in prog.recorders:
            rec.finish()
        self.evaluations += 1

        def per_member(rec):
            values = np.asarray(rec.values(), dtype=np.float64)
            agg = Aggregator(rec.agg_func)
            return np.array([agg.aggregate_1d(np.ascontiguousarray(values[self.member == p]), ignore_nan=rec.ignore_nan)
                             for p in range(self.P)])

        objectives = np.empty((self.P, len(self.objectives)))
        for j, rec in enumerate(self.objectives):
            sign = 1.0 if rec.is_objective == "minimise" else -1.0
            objectives[:, j] = sign * per_member(rec)
        cons = []
        for c in self.constraints:
            v = per_member(c)
            cons += [v, v] if c.is_double_bounded_constraint else [v]
        constraints = np.stack(cons, axis=1) if cons else np.empty((self.P, 0))
        return objectives, constraints

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None


class PlatypusPopulationWrapper:
    """Platypus front end of :class:`PopulationEvaluator` — the batched counterpart of the reference's
    ``PlatypusWrapper`` (``pywr/optimisation/platypus.py:31-154``): the same ``platypus.Problem`` (one ``Real`` per double /
    integer variable with the parameter's bounds, the same ``Constraint`` objects for lower / upper / equality / double
    bounded recorders, :60-112), but a generation is evaluated in ONE device-resident run instead of one ``Model.run()`` per
    candidate.  Platypus algorithms hand their unevaluated solutions to ``algorithm.evaluator.evaluate_all(jobs)``;
    :attr:`evaluator` is a ``platypus.Evaluator`` that packs them into populations of ``population_size`` candidates
    (the last one padded by repeating a candidate), calls :meth:`PopulationEvaluator.evaluate` and stores objectives,
    constraints, constraint violation and feasibility on every solution the way ``platypus.Problem.evaluate`` does.

        wrapper = PlatypusPopulationWrapper(make_model, population_size=256)
        algorithm = platypus.NSGAII(wrapper.problem, population_size=256, evaluator=wrapper.evaluator)
        algorithm.run(10000)

    ``platypus``: the module to build the problem with (default: ``import platypus``)."""

    def __init__(self, make_model, population_size, platypus=None, **kwargs):
        if platypus is None:
            import platypus  # noqa: PLC0415 - optional dependency, exactly like the reference's wrapper
        self.platypus = platypus
        self.batch = PopulationEvaluator(make_model, population_size, **kwargs)
        b = self.batch
        n_con = sum(2 if c.is_double_bounded_constraint else 1 for c in b.constraints)
        if len(b.variables) < 1:
            raise ValueError("At least one variable must be defined.")
        if len(b.objectives) < 1:
            raise ValueError("At least one objective must be defined.")
        self.problem = platypus.Problem(b.n_variables, len(b.objectives), n_con)
        self.problem.function = self._evaluate_one
        self.problem.wrapper = self
        ix = 0
        for var in b.variables:           # _make_variables, :60-79 (integer variables are refused by the evaluator)
            lower, upper = var.get_double_lower_bounds(), var.get_double_upper_bounds()
            for i in range(var.double_size):
                self.problem.types[ix] = platypus.Real(lower[i], upper[i])
                ix += 1
        ic = 0
        for c in b.constraints:           # _make_constraints, :81-112
            if c.is_double_bounded_constraint:
                self.problem.constraints[ic] = platypus.Constraint(">=", value=c.constraint_lower_bounds)
                self.problem.constraints[ic + 1] = platypus.Constraint("<=", value=c.constraint_upper_bounds)
                ic += 2
            elif c.is_equality_constraint:
                self.problem.constraints[ic] = platypus.Constraint("==", value=c.constraint_lower_bounds)
                ic += 1
            elif c.is_lower_bounded_constraint:
                self.problem.constraints[ic] = platypus.Constraint(">=", value=c.constraint_lower_bounds)
                ic += 1
            elif c.is_upper_bounded_constraint:
                self.problem.constraints[ic] = platypus.Constraint("<=", value=c.constraint_upper_bounds)
                ic += 1
            else:
                raise RuntimeError(f'The bounds of constraint "{c.name}" could not be identified correctly.')
        wrapper = self

        class _BatchEvaluator(platypus.Evaluator):
            def evaluate_all(self, jobs, **kw):
                jobs = list(jobs)
                wrapper.evaluate_solutions([job.solution for job in jobs])
                return jobs

        self.evaluator = _BatchEvaluator()

    def _decode(self, solution):
        return [t.decode(v) for t, v in zip(self.problem.types, solution.variables)]

    def evaluate_solutions(self, solutions):
        """Evaluates any number of platypus solutions, ``population_size`` candidates per device run."""
        P = self.batch.P
        todo = [s for s in solutions if not getattr(s, "evaluated", False)]
        for start in range(0, len(todo), P):
            chunk = todo[start: start + P]
            X = np.array([self._decode(s) for s in chunk], dtype=np.float64)
            if len(chunk) < P:
                X = np.vstack([X, np.repeat(X[-1:], P - len(chunk), axis=0)])
            objectives, constraints = self.batch.evaluate(X)
            for p, s in enumerate(chunk):
                s.objectives[:] = [float(v) for v in objectives[p]]
                s.constraints[:] = [float(v) for v in constraints[p]]
                s.constraint_violation = sum(abs(f(x)) for f, x in zip(self.problem.constraints, s.constraints))
                s.feasible = s.constraint_violation == 0.0
                s.evaluated = True

    def _evaluate_one(self, variables):
        """``problem.function`` for callers that evaluate one solution at a time (the whole batch carries that candidate)."""
        X = np.repeat(np.asarray(variables, dtype=np.float64)[None, :], self.batch.P, axis=0)
        objectives, constraints = self.batch.evaluate(X)
        if constraints.shape[1] > 0:
            return list(objectives[0]), list(constraints[0])
        return list(objectives[0])

    def close(self):
        self.batch.close()
