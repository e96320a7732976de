"""Pywr solver plug-in: ``Model(solver="b200")`` / ``"solver": {"name": "b200"}`` / ``PYWR_SOLVER=b200``.

Host side of the B200 scenario-batched LP engine.  Mirrors the interface of the reference's
``CythonGLPKSolver`` wrapper (``pywr/solvers/__init__.py:75-194``): ``setup(model)``,
``solve(model)``, ``reset()``, ``stats``, ``settings`` and the attributes the reference test-suite
reads (``use_unsafe_api``, ``set_fixed_*_once``, ``use_presolve``, ``retry_solve``,
``save_routes_flows``, ``routes``, ``routes_flows_array``).

Where the reference loops ``for scenario_index in model.scenarios.combinations:
_solve_scenario(...)`` (``cython_glpk.pyx:821-831``), this solver gathers every dynamic bound /
cost for ALL scenarios into scenario-contiguous planes of one pinned buffer (using the vectorised
``get_all_*`` getters, ``pywr/_core.pyx:566-574,618-626,681-689,1202-1238``), makes ONE engine call
per timestep, and commits the returned node flows with ``commit_all`` (``_core.pyx:485-489``).
"""

from __future__ import annotations

import time

import numpy as np

from .structure import LPStructure, build_structure

try:  # the reference only defines GLPKError when its GLPK extension was built
    from pywr.solvers import GLPKError as _BaseLPError
except Exception:  # pragma: no cover - depends on the pywr build
    _BaseLPError = Exception


try:
    from pywr.solvers import Solver as _SolverBase
except Exception:  # pragma: no cover - pywr absent: the C-ABI / engine are still usable on their own
    _SolverBase = object


class B200LPError(_BaseLPError):
    """NaN in a bound, cost or coefficient (role of ``GLPKError``, ``cython_glpk.pyx:286-335``)."""


_STATUS_TEXT = {
    1: "no primal feasible solution",
    2: "no dual feasible solution",
    3: "iteration limit exceeded",
    4: "NaN detected",
    5: "singular matrix",
}


class BatchedLPSolver(_SolverBase):
    """Engine-agnostic host logic.  Subclasses provide ``_make_engine(structure, n_scenarios)``
    returning an object with ``reset(vol0)``, ``step(days, slots, node_flow, col_flow, objective)``,
    ``status()``, ``alloc_planes(n)``, ``set_consts(consts)``, ``close()`` and ``counters()``."""

    name = "batched-lp"
    formulation = "route"

    def __init__(self, save_routes_flows=False, retry_solve=False, use_presolve=False, check_nan=True, device_run=False,
                 device_run_devices=None, device_run_arrays="eager", **kwargs):
        super().__init__()
        # device_run (JSON: "solver": {"name": "b200", "device_run": true}): Model.run() executes the whole run on the
        # device (pywr_b200.device_run.run_on_device) instead of one solve per timestep; "require" raises when the model
        # needs the host every timestep, True falls back to the per-timestep loop then.  device_run_devices / _arrays are
        # run_on_device's `devices` and `arrays`.
        self.device_run = device_run
        self.device_run_devices = device_run_devices
        self.device_run_arrays = device_run_arrays
        self.save_routes_flows = save_routes_flows
        self.retry_solve = retry_solve  # the engine always retries from the standard basis
        self.use_presolve = use_presolve
        self.check_nan = check_nan
        self.use_unsafe_api = not check_nan
        self.set_fixed_flows_once = False
        self.set_fixed_costs_once = False
        self.set_fixed_factors_once = False
        self.engine_args = kwargs
        self._engine = None
        self._structure: LPStructure | None = None
        self._model = None
        self._stats = None
        self.routes = None
        self.routes_flows_array = None
        self.last_objective = None

    # -- reference attributes ---------------------------------------------------------------
    @property
    def stats(self):
        return self._stats

    @property
    def settings(self):
        return {
            "use_presolve": self.use_presolve,
            "save_routes_flows": self.save_routes_flows,
            "retry_solve": self.retry_solve,
            "set_fixed_flows_once": self.set_fixed_flows_once,
            "set_fixed_costs_once": self.set_fixed_costs_once,
            "set_fixed_factors_once": self.set_fixed_factors_once,
            "formulation": self.formulation,
        }

    @property
    def structure(self) -> LPStructure:
        return self._structure

    def _make_engine(self, structure, n_scenarios):  # pragma: no cover - abstract
        raise NotImplementedError

    def _make_program_engine(self, structure, n_scenarios):
        """engine of a device-resident run (``device_run``); the same kind as the per-timestep one by default"""
        return self._make_engine(structure, n_scenarios)

    # -- Solver API -----------------------------------------------------------------------------
    def setup(self, model):
        """Role of ``CythonGLPKSolver.setup`` (``cython_glpk.pyx:388-755``)."""
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        s = build_structure(model, self.formulation)
        S = len(model.scenarios.combinations)
        self._model = model
        self._structure = s
        self._S = S
        self.routes = s.routes
        self._engine = self._make_engine(s, S)
        self._slots = self._engine.alloc_planes(max(s.n_slots, 1))
        self._node_flow = self._engine.alloc_planes(s.n_nodes)
        self._col_flow = self._engine.alloc_planes(s.n_cols) if self.save_routes_flows else None
        self._objective = self._engine.alloc_planes(1)
        self._has_flow = np.diff(s.flow_indptr) > 0
        self.routes_flows_array = None
        self._stats = {
            "total": 0.0, "lp_solve": 0.0, "result_update": 0.0, "constraint_update_factors": 0.0,
            "bounds_update_nonstorage": 0.0, "bounds_update_storage": 0.0, "objective_update": 0.0,
            "number_of_rows": s.n_rows, "number_of_cols": s.n_cols, "number_of_nonzero": s.nnz,
            "number_of_routes" if s.formulation == "route" else "number_of_edges": s.n_cols,
            "number_of_nodes": s.n_nodes,
        }
        self._engine.reset(None)

    def reset(self):
        """``Model.__init__`` calls this before any ``setup`` (``pywr/_model.pyx:125``)."""
        if self._engine is not None:
            self._engine.reset(None)

    def solve(self, model):
        t0 = time.perf_counter()
        s = self._structure
        ts = model.timestep
        self._gather(model)
        t1 = time.perf_counter()
        nbad = self._engine.step(float(ts.days), self._slots, self._node_flow, self._col_flow, self._objective)
        t2 = time.perf_counter()
        if nbad:
            sidx, code = self._engine.status()
            if code == 4:
                raise B200LPError(f"NaN detected in the LP data of scenario {sidx} at timestep {ts.index}")
            print("Scenario ID: {}".format(sidx))
            print("Timestep index: {}".format(ts.index))
            raise RuntimeError('Simplex solver failed with message: "{}", status: "{}".'.format(
                _STATUS_TEXT.get(code, "solver failed"), "solution is undefined"))
        # commit (cython_glpk.pyx:1091-1093): every node in sorted order; nodes that carry no
        # column (Storage parents, aggregated / virtual nodes) would receive 0 and are skipped
        node_flow = self._node_flow
        for i, node in enumerate(s.all_nodes):
            if self._has_flow[i]:
                node.commit_all(node_flow[i])
        if self.save_routes_flows:
            self.routes_flows_array = np.ascontiguousarray(self._col_flow.T)
        self.last_objective = self._objective[0].copy()
        t3 = time.perf_counter()
        st = self._stats
        st["bounds_update_nonstorage"] += t1 - t0
        st["lp_solve"] += t2 - t1
        st["result_update"] += t3 - t2
        st["total"] += t3 - t0

    # -- per-timestep gather of bounds / costs / volumes (a3-a6) ---------------------------------
    def _gather(self, model):
        from pywr.parameters import Parameter

        s = self._structure
        slots = self._slots
        consts_dirty = False
        for b in s.bindings:
            kind = b.kind
            obj = b.obj
            if b.slot >= 0:
                out = slots[b.slot]
                if kind == "volume":
                    out[:] = obj._volume
                elif kind == "active":
                    out[:] = 1.0 if obj.active else 0.0
                elif kind == "factor_norm":
                    f = obj.factors
                    np.divide(np.asarray(f[0].get_all_values()), np.asarray(f[b.index].get_all_values()), out=out)
                elif b.loop:
                    getter = getattr(obj, "get_" + kind)
                    for si in model.scenarios.combinations:
                        out[si.global_id] = getter(si)
                else:
                    getattr(obj, "get_all_" + kind)(out)
            elif kind == "factor_check":
                # constant factors are part of the matrix: a ConstantParameter factor that was given a new value between
                # runs (optimisation variables, value assignment) must reach the LP like it does in the reference
                if float(np.asarray(obj.get_factors_norm(None))[b.index]) != b.baked:
                    self.setup(model)
                    return self._gather(model)
            else:
                value = _current_value(obj, kind)
                if isinstance(value, Parameter):
                    # an attribute that was a literal at setup became a Parameter: rebuild
                    self.setup(model)
                    return self._gather(model)
                if value != s.consts[b.const] and not (value != value and s.consts[b.const] != s.consts[b.const]):
                    s.consts[b.const] = value
                    consts_dirty = True
        if consts_dirty:
            self._engine.set_consts(s.consts)
        if self.check_nan and s.n_slots and np.isnan(slots[: s.n_slots]).any():
            plane, sidx = np.argwhere(np.isnan(slots[: s.n_slots]))[0]
            b = next(b for b in s.bindings if b.slot == plane)
            raise B200LPError(
                f"NaN detected in {b.kind} of {getattr(b.obj, 'name', b.obj)!r} for scenario {sidx}")
        if self.check_nan and np.isnan(s.consts).any():
            raise B200LPError("NaN detected in a constant bound or cost")

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def _current_value(obj, kind):
    from pywr._core import StorageInput, StorageOutput

    if kind == "cost" and isinstance(obj, (StorageInput, StorageOutput)):
        c = obj.parent.cost
        if isinstance(c, (int, float)):
            return -c if isinstance(obj, StorageInput) else c
        return c
    return getattr(obj, kind)


class B200Solver(BatchedLPSolver):
    """The CUDA (sm_100a) engine behind Pywr's solver API.  No CPU fallback: constructing the
    engine raises if the extension library or a CUDA device is missing."""

    name = "b200"
    formulation = "route"

    def _make_engine(self, structure, n_scenarios):
        from .engine import CudaEngine

        return CudaEngine(structure, n_scenarios, **self.engine_args)

    def _make_program_engine(self, structure, n_scenarios):
        from .engine import CudaEngine, ShardedEngine

        if self.device_run_devices is not None:
            return ShardedEngine(structure, n_scenarios, devices=self.device_run_devices, for_program=True)
        return CudaEngine(structure, n_scenarios, for_program=True, **self.engine_args)


class B200EdgeSolver(B200Solver):
    """Edge formulation (role of ``CythonGLPKEdgeSolver``, ``cython_glpk.pyx:1098-1918``)."""

    name = "b200-edge"
    formulation = "edge"


# what Model.run() did for models whose solver asked for device_run: ("device", None) or ("host", reason) per call
device_run_log = []


def _install_device_run():
    """``Model.run()`` (``pywr/_model.pyx:637-682``) is a loop of ``solver.solve`` calls that a solver cannot take over from
    the inside.  For models whose solver was created with ``device_run`` the method is routed to the device-resident run,
    which returns the same :class:`ModelResult`; every other model runs the reference's own method untouched."""
    import time

    import pywr
    from pywr import _model

    if getattr(_model.Model.run, "_b200_device_run", False):
        return
    original = _model.Model.run

    def run(self):
        solver = self.solver
        mode = getattr(solver, "device_run", False)
        if not mode:
            return original(self)
        from pywr.parameters import Parameter
        from pywr.recorders import Recorder

        from .device_run import run_on_device
        from .program import Unsupported

        # bare components have per-timestep hooks: decided before anything is set up or reset, so that the reference's run()
        # sees the model exactly as it would have
        if mode != "require" and any(not isinstance(c, (Parameter, Recorder)) for c in self.components):
            device_run_log.append(("host", "bare component"))
            return original(self)
        t0 = time.time()
        try:
            res = run_on_device(self, formulation=solver.formulation, arrays=solver.device_run_arrays,
                                engine_factory=solver._make_program_engine)
        except Unsupported as exc:
            if mode == "require":
                raise
            device_run_log.append(("host", str(exc)))
            return original(self)
        device_run_log.append(("device", None))
        total = time.time() - t0
        timestep = self.timestep
        time_taken = res.time_compile + res.time_run + res.time_readout
        return _model.ModelResult(
            num_scenarios=res.num_scenarios, timestep=timestep, time_taken=time_taken, time_taken_before=0.0,
            time_taken_after=0.0, time_taken_with_overhead=total,
            speed=(timestep.index * res.num_scenarios) / time_taken if time_taken > 0 else float("nan"),
            solver_name=solver.name, solver_settings=dict(solver.settings, device_run=True),
            solver_stats=dict(solver.stats or {}, device_run_s=res.time_run, compile_s=res.time_compile, readout_s=res.time_readout,
                              pivots=res.counters.get("pivots")),
            version=pywr.__version__)

    run._b200_device_run = True
    run.__doc__ = original.__doc__
    _model.Model.run = run


def register():
    """Append the solvers to ``pywr.solvers.solver_registry`` (``pywr/solvers/__init__.py:14``)."""
    import pywr.solvers as ps

    _install_device_run()

    for cls in (B200Solver, B200EdgeSolver):
        if not any(getattr(c, "name", None) == cls.name for c in ps.solver_registry):
            # make the classes real subclasses of pywr's Solver for isinstance checks
            ps.solver_registry.append(cls)
    return ps.solver_registry
