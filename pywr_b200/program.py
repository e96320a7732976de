"""Device-resident run: compile a Pywr model into a *program* the engine executes for all
timesteps without host round trips (north_star: "flow commit, storage
mass balance and the built-in recorders' per-scenario accumulation stay on device, so results reach
the host only at recorder read-out").

The compiler walks exactly the objects the reference's run loop touches each timestep
(``Model.before/solve/after``, ``pywr/_model.pyx:626-631,825-860``):

* every ``Parameter`` that feeds a bound, cost or capacity of the LP (the ``get_min_flow`` /
  ``get_max_flow`` / ``get_cost`` / ``get_max_volume`` getters of ``pywr/_core.pyx``) is turned
  into device ops (``include/b200lp.h`` ``B200LP_OP_*``) or, when it does not depend on model
  state, expanded once into a table ``[timesteps][scenario members]`` that is uploaded to HBM;
* ``Storage.after`` / ``VirtualStorage.after`` become the storage linear forms ``stflow_*``;
* the built-in recorders named by north_star become ``B200LP_REC_*`` outputs.

Anything outside that set raises :class:`Unsupported`; the caller then uses the per-timestep
solver path (``pywr_b200.solver.B200Solver`` under ``Model.run``), which handles every model.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Optional,  Any, Dict, List

import numpy as np

from .structure import LPStructure, build_structure, REF_STATE, ST_KIND_VIRTUAL

OP_TABLE, OP_AGG, OP_CONTROL_CURVE, OP_CC_INDEX, OP_INDEXED, OP_NEG, OP_MAX0, OP_MIN0, OP_DIV = range(1, 10)
OP_CC_INTERP, OP_CC_PIECEWISE, OP_THRESHOLD, OP_INTERP, OP_LAG, OP_STATE = 11, 12, 13, 14, 15, 16
PRED = {"LT": 0, "GT": 1, "EQ": 2, "LE": 3, "GE": 4}
OP_STORAGE_RESET = 10
PROG_AGG_ONLY = 1
# virtual storages whose reset / activation follows the calendar only (pywr/nodes.py:596-757)
SCHEDULED_STORAGES = ("AnnualVirtualStorage", "SeasonalVirtualStorage", "MonthlyVirtualStorage")
AGG = {"sum": 0, "product": 1, "min": 2, "max": 3, "mean": 4}
COL_NONE, COL_SCENARIO = -1, -2
(REC_NODE_FLOW, REC_NODE_DEFICIT, REC_NODE_SUPPLIED, REC_NODE_CURTAIL, REC_STORAGE_VOLUME, REC_STORAGE_PC,
 REC_TOTAL_DEFICIT, REC_TOTAL_FLOW, REC_MEAN_FLOW, REC_DEFICIT_FREQ, REC_MIN_VOLUME) = range(1, 12)
ARRAY_KINDS = {REC_NODE_FLOW, REC_NODE_DEFICIT, REC_NODE_SUPPLIED, REC_NODE_CURTAIL, REC_STORAGE_VOLUME, REC_STORAGE_PC}


class Unsupported(Exception):
    """The model uses something the device-resident run does not cover (yet)."""


class b200lp_program(ctypes.Structure):
    _i32p = ctypes.POINTER(ctypes.c_int32)
    _i64p = ctypes.POINTER(ctypes.c_int64)
    _f64p = ctypes.POINTER(ctypes.c_double)
    _fields_ = [
        ("n_steps", ctypes.c_int32), ("n_ops", ctypes.c_int32), ("n_tables", ctypes.c_int32),
        ("n_axes", ctypes.c_int32), ("n_recorders", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("ts_days", _f64p),
        ("op_kind", _i32p), ("op_dst", _i32p), ("op_arg_ptr", _i32p), ("op_args", _i32p),
        ("table_rows", _i32p), ("table_cols", _i32p), ("table_axis", _i32p), ("table_offset", _i64p),
        ("table_data", _f64p),
        ("scen_index", _i32p),
        ("stflow_indptr", _i32p), ("stflow_col", _i32p), ("stflow_w", _f64p),
        ("initial_volume", _f64p), ("initial_pc", _f64p),
        ("rec_kind", _i32p), ("rec_src", _i32p), ("rec_ref", _i32p), ("rec_factor", _f64p),
        ("rec_row", _i32p), ("rec_indptr", _i32p), ("rec_col", _i32p), ("rec_w", _f64p),
        ("st_roll_len", _i32p), ("st_roll_init", _f64p),
    ]


_P_I32 = ("op_kind op_dst op_arg_ptr op_args table_rows table_cols table_axis scen_index stflow_indptr "
          "stflow_col rec_kind rec_src rec_ref rec_row rec_indptr rec_col").split()
_P_I64 = ["table_offset"]
_P_F64 = "ts_days table_data stflow_w initial_volume initial_pc rec_factor rec_w".split()


@dataclass
class Program:
    n_steps: int
    n_axes: int
    ts_days: np.ndarray
    op_kind: np.ndarray
    op_dst: np.ndarray
    op_arg_ptr: np.ndarray
    op_args: np.ndarray
    table_rows: np.ndarray
    table_cols: np.ndarray
    table_axis: np.ndarray
    table_offset: np.ndarray
    table_data: np.ndarray
    scen_index: np.ndarray
    stflow_indptr: np.ndarray
    stflow_col: np.ndarray
    stflow_w: np.ndarray
    initial_volume: np.ndarray
    initial_pc: np.ndarray
    rec_kind: np.ndarray
    rec_src: np.ndarray
    rec_ref: np.ndarray
    rec_factor: np.ndarray
    rec_row: np.ndarray
    rec_indptr: np.ndarray
    rec_col: np.ndarray
    rec_w: np.ndarray
    recorders: List[Any] = field(default_factory=list, repr=False)  # host objects, same order as rec_*
    rec_names: List[str] = field(default_factory=list)
    host_recorders: List[Any] = field(default_factory=list, repr=False)  # kept alive by the host mirror protocol
    variable_tables: Dict[int, int] = field(default_factory=dict, repr=False)  # id(variable parameter) -> table
    months: Optional[np.ndarray] = None        # [T] calendar month of every timestep (population batches)
    st_roll_len: Optional[np.ndarray] = None   # [n_storage] RollingVirtualStorage: timesteps - 1 (0: not rolling)
    st_roll_init: Optional[np.ndarray] = None  # [n_storage] its _initial_utilisation
    flags: int = 0                             # B200LP_PROG_*: 1 = array recorders keep their temporal aggregates only

    @property
    def n_ops(self):
        return int(self.op_kind.shape[0])

    @property
    def n_tables(self):
        return int(self.table_rows.shape[0])

    @property
    def n_recorders(self):
        return int(self.rec_kind.shape[0])

    @property
    def has_lag(self):
        """some op reads a recorder's sample of the previous timestep: the recorder arrays are state (no PROG_AGG_ONLY)"""
        return bool(np.any(np.asarray(self.op_kind) == OP_LAG))

    def rec_shape(self, r, S):
        return (self.n_steps, S) if int(self.rec_kind[r]) in ARRAY_KINDS else (S,)

    def to_c(self) -> b200lp_program:
        cp = b200lp_program()
        cp.n_steps, cp.n_ops, cp.n_tables = self.n_steps, self.n_ops, self.n_tables
        cp.n_axes, cp.n_recorders = self.n_axes, self.n_recorders
        cp.flags = int(self.flags)
        keep = []
        for names, dt, ct in ((_P_I32, np.int32, ctypes.c_int32), (_P_I64, np.int64, ctypes.c_int64),
                              (_P_F64, np.float64, ctypes.c_double)):
            for name in names:
                arr = np.ascontiguousarray(getattr(self, name), dtype=dt).reshape(-1)
                if arr.size == 0:
                    arr = np.zeros(1, dt)
                keep.append(arr)
                setattr(cp, name, arr.ctypes.data_as(ctypes.POINTER(ct)))
        n_st = max(len(self.stflow_indptr) - 1, 1)
        rl = np.zeros(n_st, np.int32) if self.st_roll_len is None else np.ascontiguousarray(self.st_roll_len, np.int32)
        ri = np.zeros(n_st, np.float64) if self.st_roll_init is None else np.ascontiguousarray(self.st_roll_init, np.float64)
        keep += [rl, ri]
        cp.st_roll_len = rl.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        cp.st_roll_init = ri.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        cp._keepalive = keep
        return cp

    def shard(self, lo, hi, S):
        """The program of scenarios [lo, hi) of a batch of S (contiguous ``global_id`` block, pywr/_core.pyx:210-240):
        per-scenario tables keep their columns lo..hi-1, scenario indices and initial states are sliced; tables
        indexed by a scenario axis' members, ops and recorders are shared.  Scenarios are independent
        (cython_glpk.pyx:827-828), so running the shards on different devices is the same computation."""
        import copy

        q = copy.copy(self)
        n = hi - lo
        rows, cols, offs, chunks, pos = [], [], [], [], 0
        for t in range(self.n_tables):
            r0, c0 = int(self.table_rows[t]), int(self.table_cols[t])
            data = self.table_data[int(self.table_offset[t]): int(self.table_offset[t]) + r0 * c0].reshape(r0, c0)
            if int(self.table_axis[t]) == COL_SCENARIO:
                data = data[:, lo:hi]
            rows.append(data.shape[0]); cols.append(data.shape[1]); offs.append(pos)
            chunks.append(np.ascontiguousarray(data).reshape(-1)); pos += data.size
        q.table_rows, q.table_cols = np.array(rows, np.int32), np.array(cols, np.int32)
        q.table_offset = np.array(offs, np.int64)
        q.table_data = np.concatenate(chunks) if chunks else np.zeros(0)
        if self.n_axes > 0:
            q.scen_index = np.ascontiguousarray(np.asarray(self.scen_index, np.int32).reshape(self.n_axes, S)[:, lo:hi]).reshape(-1)
        nst = max(len(self.stflow_indptr) - 1, 0)
        if nst > 0:
            q.initial_volume = np.ascontiguousarray(np.asarray(self.initial_volume).reshape(nst, S)[:, lo:hi]).reshape(-1)
            q.initial_pc = np.ascontiguousarray(np.asarray(self.initial_pc).reshape(nst, S)[:, lo:hi]).reshape(-1)
        assert q.initial_volume.size == nst * n
        return q

    def save(self, path):
        data = {name: np.asarray(getattr(self, name)) for name in _P_I32 + _P_I64 + _P_F64}
        data["meta"] = np.array([self.n_steps, self.n_axes], np.int64)
        if self.st_roll_len is not None:
            data["st_roll_len"], data["st_roll_init"] = np.asarray(self.st_roll_len, np.int32), np.asarray(self.st_roll_init, np.float64)
        data["rec_names"] = np.array(self.rec_names if self.rec_names else [""] * self.n_recorders)
        np.savez_compressed(path, **data)

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        n_steps, n_axes = (int(v) for v in z["meta"])
        kw = {n: z[n].astype(np.int32) for n in _P_I32}
        kw.update({n: z[n].astype(np.int64) for n in _P_I64})
        kw.update({n: z[n].astype(np.float64) for n in _P_F64})
        names = [str(x) for x in z["rec_names"]] if "rec_names" in z.files else []
        if "st_roll_len" in z.files:
            kw["st_roll_len"], kw["st_roll_init"] = z["st_roll_len"].astype(np.int32), z["st_roll_init"].astype(np.float64)
        return cls(n_steps=n_steps, n_axes=n_axes, rec_names=names, **kw)


# ----------------------------------------------------------------------------------------------
# compiler
# ----------------------------------------------------------------------------------------------
_VALUE_LEAVES = {  # state- and scenario-independent: sampled through their public ``value(ts, si)``
    "MonthlyProfileParameter", "DailyProfileParameter", "WeeklyProfileParameter",
    "UniformDrawdownProfileParameter", "RbfProfileParameter", "AnnualHarmonicSeriesParameter", "DiscountFactorParameter",
}
_THRESHOLDS = {"ParameterThresholdParameter", "StorageThresholdParameter", "CurrentYearThresholdParameter",
               "CurrentOrdinalDayThresholdParameter", "NodeThresholdParameter", "RecorderThresholdParameter"}
_SCENARIO_PROFILES = {
    "ScenarioMonthlyProfileParameter", "ScenarioDailyProfileParameter", "ScenarioWeeklyProfileParameter",
    "ArrayIndexedScenarioMonthlyFactorsParameter",
}


class _Compiler:
    def __init__(self, model, structure: LPStructure):
        from pywr import _core

        self.model = model
        self.s = structure
        self.S = len(model.scenarios.combinations)
        ts = model.timestepper
        self.T = len(ts)
        self._timesteps = None          # Timestep objects are built on first use (10,957 of them cost half a second)
        self.days = np.asarray(ts._deltas[: self.T], dtype=np.float64)
        self.months = np.asarray(ts._periods.month[: self.T], dtype=np.int32)
        self.combinations = list(model.scenarios.combinations)
        self.n_axes = len(model.scenarios.scenarios)
        self.scen_index = np.zeros((max(self.n_axes, 1), self.S), np.int32)
        if self.n_axes and self.S:
            idx = np.array([np.asarray(si.indices) for si in self.combinations], dtype=np.int32).reshape(self.S, -1)
            gid = np.fromiter((si.global_id for si in self.combinations), dtype=np.int64, count=self.S)
            self.scen_index[: self.n_axes, gid] = idx[:, : self.n_axes].T
        self.consts = list(structure.consts)
        self.n_slots = 0
        self.ops = []      # (kind, dst, args)
        self.tables = []   # (values[rows, cols], axis)
        self.memo: Dict[int, int] = {}
        self.storage_index = {id(st): k for k, st in enumerate(structure.storages)}
        # optimisation batch: variable parameters become tables over this scenario axis, one column
        # per candidate solution, so that a whole population is evaluated in one run (optimisation.PopulationEvaluator)
        self.population_axis = None
        self.variable_tables: Dict[int, int] = {}   # id(parameter) -> table index
        self.extra_storages = []                    # AggregatedStorage nodes: device storages without an LP row

    @property
    def timesteps(self):
        if self._timesteps is None:
            from pywr import _core

            ts = self.model.timestepper
            periods, deltas = list(ts._periods[: self.T]), list(ts._deltas[: self.T])
            self._timesteps = [_core.Timestep(periods[i], i, deltas[i]) for i in range(self.T)]
        return self._timesteps

    def monthly_rows(self, sample):
        """[T, k] table of a quantity that depends on the calendar month only: `sample(ts)` (a row of k numbers) is taken
        at the first timestep of every month that occurs and the rows are expanded by the month of each timestep."""
        from pywr import _core

        ts = self.model.timestepper
        first = {}
        for i, mth in enumerate(self.months):
            if int(mth) not in first:
                first[int(mth)] = i
                if len(first) == 12:
                    break
        rows = {mth: np.asarray(sample(_core.Timestep(ts._periods[i], i, ts._deltas[i])), dtype=np.float64) for mth, i in first.items()}
        k = len(next(iter(rows.values())))
        table = np.zeros((13, k))
        for mth, row in rows.items():
            table[mth] = row
        return table[self.months]

    # -- small helpers ---------------------------------------------------------------------
    def const(self, v) -> int:
        self.consts.append(float(v))
        return -len(self.consts)

    def new_slot(self) -> int:
        self.n_slots += 1
        return self.n_slots - 1

    def emit(self, kind, args) -> int:
        dst = self.new_slot()
        self.ops.append((kind, dst, [int(a) for a in args]))
        return dst

    def table(self, values, axis, offset=0) -> int:
        values = np.ascontiguousarray(values, dtype=np.float64)
        assert values.ndim == 2
        if values.shape == (1, 1) and offset == 0:
            return self.const(values[0, 0])
        self.tables.append((values, COL_NONE if axis is None else int(axis)))
        return self.emit(OP_TABLE, [len(self.tables) - 1, offset])

    def representative(self, axis, member):
        """A ScenarioIndex whose index on ``axis`` is ``member`` (other axes at 0)."""
        from pywr._core import ScenarioIndex

        idx = np.zeros(max(self.n_axes, 1), np.int32)
        if axis is not None:
            idx[axis] = member
        return ScenarioIndex(0, idx)

    # -- parameters -> refs ----------------------------------------------------------------
    def value_ref(self, value) -> int:
        from pywr.parameters import Parameter

        if isinstance(value, Parameter):
            return self.param(value)
        return self.const(value)

    def param(self, p) -> int:
        key = id(p)
        if key in self.memo:
            return self.memo[key]
        ref = self._param(p)
        self.memo[key] = ref
        return ref

    def _param(self, p) -> int:
        from .bridge import _bridge

        name = type(p).__name__
        if self.population_axis is not None and getattr(p, "is_variable", False):
            if name not in ("ConstantParameter", "MonthlyProfileParameter"):
                raise Unsupported(f"variable parameter of type {name} in a population batch")
            P = self.model.scenarios.scenarios[self.population_axis].size
            si = self.representative(None, 0)
            if name == "ConstantParameter":
                rows = np.full((1, P), p.get_constant_value(), dtype=np.float64)
            else:
                rows = np.repeat(self.monthly_rows(lambda ts: [p.value(ts, si)]), P, axis=1)
            self.tables.append((np.ascontiguousarray(rows, dtype=np.float64), int(self.population_axis)))
            self.variable_tables[id(p)] = len(self.tables) - 1
            return self.emit(OP_TABLE, [len(self.tables) - 1, 0])
        if name == "ConstantParameter" or (getattr(p, "is_constant", False) and name != "ConstantScenarioParameter"):
            return self.const(p.get_constant_value())
        info = _bridge.parameter_arrays(p)
        if info is not None:
            values, axis = info["values"], info["axis"]
            if info["col_map"] is not None:  # sliced / user-combination data: member -> loaded column
                cmap = info["col_map"]
                full = np.full((values.shape[0], len(cmap)), np.nan)
                ok = cmap < values.shape[1]
                full[:, ok] = values[:, cmap[ok]]
                values = full
            if values.shape[0] > self.T and name in ("ArrayIndexedParameter", "ArrayIndexedScenarioParameter"):
                values = values[: self.T]          # values[ts.index]: the rows past the run are never read
            if values.shape[0] not in (1, self.T):
                raise Unsupported(f"{name}: {values.shape[0]} rows for {self.T} timesteps")
            return self.table(values, axis, int(info["offset"]))
        if name in _VALUE_LEAVES:
            si = self.representative(None, 0)
            if name == "MonthlyProfileParameter":   # value(ts) = values[ts.month - 1]: 12 samples instead of T
                return self.table(self.monthly_rows(lambda ts: [p.value(ts, si)]), None)
            return self.table(np.array([[p.value(ts, si)] for ts in self.timesteps]), None)
        if name in _SCENARIO_PROFILES:
            axis = _bridge.scenario_axis(p)
            size = self.model.scenarios.scenarios[axis].size
            sis = [self.representative(axis, c) for c in range(size)]
            if name == "ScenarioMonthlyProfileParameter":
                return self.table(self.monthly_rows(lambda ts: [p.value(ts, si) for si in sis]), axis)
            return self.table(np.array([[p.value(ts, si) for si in sis] for ts in self.timesteps]), axis)
        if name == "AggregatedParameter":
            func = p.agg_func
            if not isinstance(func, str) or func not in AGG:
                raise Unsupported(f"AggregatedParameter agg_func {func!r}")
            refs = [self.param(c) for c in p.parameters]
            return self.emit(OP_AGG, [AGG[func], len(refs)] + refs)
        if name == "ControlCurveParameter":
            k = self.storage_of(p.storage_node)
            curves = [self.param(c) for c in p.control_curves]
            if p.parameters is not None:
                vals = [self.param(c) for c in p.parameters]
            else:
                vals = [self.const(v) for v in np.asarray(p.values)]
            if len(vals) != len(curves) + 1:
                raise Unsupported("ControlCurveParameter: len(values) != len(control_curves)+1")
            return self.emit(OP_CONTROL_CURVE, [k, len(curves)] + curves + vals)
        if name == "ControlCurveInterpolatedParameter":   # pywr/parameters/_control_curves.pyx:98-228
            k = self.storage_of(p.storage_node)
            curves = [self.param(c) for c in p.control_curves]
            if p.parameters is not None:
                vals = [self.param(c) for c in p.parameters]
            else:
                vals = [self.const(v) for v in np.asarray(p.values)]
            if len(vals) != len(curves) + 2:
                raise Unsupported("ControlCurveInterpolatedParameter: len(values) != len(control_curves)+2")
            return self.emit(OP_CC_INTERP, [k, len(curves)] + curves + vals)
        if name == "ControlCurvePiecewiseInterpolatedParameter":   # :231-301
            k = self.storage_of(p.storage_node)
            curves = [self.param(c) for c in p.control_curves]
            values = np.asarray(p.values, dtype=np.float64)
            if values.shape != (len(curves) + 1, 2):
                raise Unsupported("ControlCurvePiecewiseInterpolatedParameter: values must be [len(control_curves)+1, 2]")
            upper = [self.const(v) for v in values[:, 0]]
            lower = [self.const(v) for v in values[:, 1]]
            return self.emit(OP_CC_PIECEWISE, [k, len(curves)] + curves + upper + lower + [self.const(p.minimum), self.const(p.maximum)])
        if name == "ControlCurveIndexParameter":
            k = self.storage_of(p.storage_node)
            curves = [self.param(c) for c in p.control_curves]
            return self.emit(OP_CC_INDEX, [k, len(curves)] + curves)
        if name == "IndexedArrayParameter":
            idx = self.index_param(p.index_parameter)
            refs = [self.param(c) for c in p.params]
            return self.emit(OP_INDEXED, [idx, len(refs)] + refs)
        if name == "NegativeParameter":
            return self.emit(OP_NEG, [self.param(p.parameter)])
        if name == "MaxParameter":
            return self.emit(OP_MAX0, [self.param(p.parameter), self.const(p.threshold)])
        if name == "MinParameter":
            return self.emit(OP_MIN0, [self.param(p.parameter), self.const(p.threshold)])
        if name == "NegativeMaxParameter":
            return self.emit(OP_MAX0, [self.emit(OP_NEG, [self.param(p.parameter)]), self.const(p.threshold)])
        if name == "NegativeMinParameter":
            return self.emit(OP_MIN0, [self.emit(OP_NEG, [self.param(p.parameter)]), self.const(p.threshold)])
        if name == "FlowParameter":                     # pywr/parameters/_parameters.pyx:1959-2008: the node's flow of the day before
            return self.prev_flow(p.node, float(p.initial_value))
        if name == "StorageParameter":                  # :2011-2044: the storage's volume / proportional volume as they stand
            return self.emit(OP_STATE, [2 if p.use_proportional_volume else 1, self.storage_of(p.storage_node)])
        if name == "Polynomial1DParameter":             # pywr/parameters/_polynomial.pyx:5-88: sum c_i x^i of a parameter / a storage
            info = _bridge.polynomial1d_info(p)
            if info["parameter"] is not None:
                x = self.param(info["parameter"])
            elif info["storage_node"] is not None:
                x = self.emit(OP_STATE, [3 if info["use_proportional_volume"] else 1, self.storage_of(info["storage_node"])])
            else:
                raise Unsupported("Polynomial1DParameter of a node's flow of the same timestep")
            x = self.emit(OP_AGG, [AGG["sum"], 2, self.emit(OP_AGG, [AGG["product"], 2, x, self.const(info["scale"])]), self.const(info["offset"])])
            terms, power = [], None
            for i, c in enumerate(info["coefficients"]):
                if i == 0:
                    terms.append(self.const(c))              # c_0 * x**0
                    continue
                power = x if i == 1 else self.emit(OP_AGG, [AGG["product"], 2, power, x])
                terms.append(self.emit(OP_AGG, [AGG["product"], 2, self.const(c), power]))
            return self.emit(OP_AGG, [AGG["sum"], len(terms)] + terms)
        if name == "Polynomial2DStorageParameter":      # _polynomial.pyx:90-175: sum c_ij x^i y^j of a storage and a parameter
            info = _bridge.polynomial2d_info(p)
            x = self.emit(OP_STATE, [3 if info["use_proportional_volume"] else 1, self.storage_of(info["storage_node"])])
            x = self.emit(OP_AGG, [AGG["sum"], 2, self.emit(OP_AGG, [AGG["product"], 2, self.const(info["storage_scale"]), x]),
                                   self.const(info["storage_offset"])])
            y = self.emit(OP_AGG, [AGG["sum"], 2, self.emit(OP_AGG, [AGG["product"], 2, self.const(info["parameter_scale"]),
                                                                      self.param(info["parameter"])]), self.const(info["parameter_offset"])])
            coef = info["coefficients"]

            def powers(base, n):
                out = [self.const(1.0)]
                for i in range(1, n):
                    out.append(base if i == 1 else self.emit(OP_AGG, [AGG["product"], 2, out[-1], base]))
                return out

            xp, yp = powers(x, coef.shape[0]), powers(y, coef.shape[1])
            terms = [self.emit(OP_AGG, [AGG["product"], 3, self.const(coef[i, j]), xp[i], yp[j]])
                     for i in range(coef.shape[0]) for j in range(coef.shape[1])]
            return self.emit(OP_AGG, [AGG["sum"], len(terms)] + terms)
        if name == "KeatingStreamFlowParameter":        # pywr/parameters/groundwater.py:5-62: sum over the stream levels below the
            level = self.value_ref(p.storage_node.level)     # aquifer's level of 2 T C (level - stream level)
            terms = []
            for L, Tn in zip(np.asarray(p.levels, dtype=np.float64), np.asarray(p.transmissivity, dtype=np.float64)):
                k = 2 * float(Tn) * float(p.coefficient)
                diff = self.emit(OP_AGG, [AGG["sum"], 2, level, self.const(-float(L))])
                term = self.emit(OP_AGG, [AGG["product"], 2, self.const(k), diff])
                terms.append(self.emit(OP_THRESHOLD, [0, level, PRED["GT"], self.const(float(L)), self.const(0.0), term]))
            return self.emit(OP_AGG, [AGG["sum"], len(terms)] + terms) if terms else self.const(0.0)
        if name == "ScenarioWrapperParameter":          # pywr/parameters/parameters.py:359-404: one child parameter per scenario member
            axis = self.model.scenarios.get_scenario_index(p.scenario)
            size = p.scenario.size
            member = self.table(np.arange(size, dtype=np.float64).reshape(1, -1), axis)
            return self.emit(OP_INDEXED, [member, size] + [self.param(c) for c in p.parameters])
        if name == "DeficitParameter":                  # :1910-1957: yesterday's max_flow - flow, 0 on the first timestep
            return self.prev_flow(p.node, 0.0, 1, deficit=True)
        if name in ("FlowDelayParameter", "RollingMeanFlowNodeParameter"):   # :2105-2163, :2191-2264
            n = int(p.timesteps)
            if int(p.days) > 0:
                delta = self.model.timestepper.delta
                delta = int(getattr(delta, "days", delta))
                n = int(p.days) // delta
            if n < 1 or n > 64:
                raise Unsupported(f"{name} over {n} timesteps")
            if name == "FlowDelayParameter":
                return self.prev_flow(p.node, float(p.initial_flow), n)
            # mean of the last min(t, n) flows; the initial flow on the first timestep
            lags = [self.prev_flow(p.node, 0.0, j) for j in range(1, n + 1)]
            total = self.emit(OP_AGG, [AGG["sum"], n] + lags)
            count = self.table(np.array([[float(min(max(i, 1), n))] for i in range(self.T)]), None)
            mean = self.emit(OP_DIV, [total, count])
            first = self.table(np.array([[1.0 if i == 0 else 0.0] for i in range(self.T)]), None)
            return self.emit(OP_THRESHOLD, [0, first, PRED["EQ"], self.const(1.0), mean, self.const(float(p.initial_flow))])
        if name == "DivisionParameter":                 # pywr/parameters/_parameters.pyx:1688-1739
            return self.emit(OP_DIV, [self.param(p.numerator), self.param(p.denominator)])
        if name == "OffsetParameter":                   # :1858-1907: parameter + offset
            return self.emit(OP_AGG, [AGG["sum"], 2, self.param(p.parameter), self.const(p.offset)])
        if name == "ScaledProfileParameter":            # pywr/parameters/parameters.py:102-121: scale * profile
            return self.emit(OP_AGG, [AGG["product"], 2, self.const(p.scale), self.param(p.profile)])
        if name == "ConstantScenarioIndexParameter":    # :1228-1260 (value() of an IndexParameter is its index)
            return self.index_param(p)
        if name in ("AggregatedIndexParameter", "MultipleThresholdParameterIndexParameter", "MultipleThresholdIndexParameter"):
            return self.index_param(p)
        if name in ("InterpolatedParameter", "InterpolatedVolumeParameter"):   # pywr/parameters/parameters.py:124-242
            kw = dict(p.interp_kwargs)
            if kw.pop("kind", "linear") != "linear":
                raise Unsupported(f"{name}: only linear interpolation runs on the device")
            bounds_error = kw.pop("bounds_error", True)
            fill = kw.pop("fill_value", np.nan)
            if kw or isinstance(fill, str):
                raise Unsupported(f"{name}: interp_kwargs {sorted(kw) or fill!r}")
            xs, ys = np.asarray(p.x, dtype=np.float64).reshape(-1), np.asarray(p.y, dtype=np.float64)
            if ys.ndim != 1 or len(xs) != len(ys) or len(xs) < 2:
                raise Unsupported(f"{name}: x / y must be two 1-D arrays of the same length")
            order = np.argsort(xs, kind="mergesort")     # interp1d sorts its knots (assume_sorted=False)
            xs, ys = xs[order], ys[order]
            if bounds_error:                             # the reference raises ValueError: here the scenario fails with ST_NAN
                below = above = np.nan
            else:
                below, above = (fill if isinstance(fill, tuple) else (fill, fill))
            src = [1, self.storage_of(p._node)] if name == "InterpolatedVolumeParameter" else [0, self.param(p.parameter)]
            return self.emit(OP_INTERP, src + [len(xs)] + [self.const(v) for v in xs] + [self.const(v) for v in ys]
                             + [self.const(below), self.const(above)])
        if name in _THRESHOLDS:                         # pywr/parameters/_thresholds.pyx
            v = _bridge.threshold_values(p)
            if v is None:
                raise Unsupported(f"{name} without values used as a value")
            return self.threshold(p, self.const(v[0]), self.const(v[1]))
        raise Unsupported(f"parameter type {name} is not available on the device")

    def index_param(self, p) -> int:
        """ref of an IndexParameter's INDEX (as a double), for IndexedArrayParameter / AggregatedIndexParameter"""
        from .bridge import _bridge

        name = type(p).__name__
        key = ("index", id(p))
        if key in self.memo:
            return self.memo[key]
        if name in _THRESHOLDS:
            ref = self.threshold(p, self.const(0.0), self.const(1.0))
        elif name == "ConstantScenarioIndexParameter":
            axis = _bridge.scenario_axis(p)
            size = self.model.scenarios.scenarios[axis].size
            ts0 = self.timesteps[0]
            ref = self.table(np.array([[float(p.index(ts0, self.representative(axis, c))) for c in range(size)]]), axis)
        elif name == "MultipleThresholdParameterIndexParameter":   # _thresholds.pyx:242-292: first threshold the value is not below
            x = self.param(p.parameter)
            thresholds = _bridge.multiple_thresholds(p)
            ref = self.const(float(len(thresholds)))
            for j in range(len(thresholds) - 1, -1, -1):
                ref = self.emit(OP_THRESHOLD, [0, x, PRED["GE"], self.param(thresholds[j]), ref, self.const(float(j))])
        elif name == "MultipleThresholdIndexParameter":            # _thresholds.pyx:190-240: the same on yesterday's flow of a node
            x = self.prev_flow(p.node)
            thresholds = _bridge.multiple_thresholds(p)
            ref = self.const(float(len(thresholds)))
            for j in range(len(thresholds) - 1, -1, -1):
                ref = self.emit(OP_THRESHOLD, [0, x, PRED["GE"], self.param(thresholds[j]), ref, self.const(float(j))])
        elif name == "AggregatedIndexParameter":
            func = p.agg_func
            if not isinstance(func, str) or func not in ("sum", "min", "max", "product"):
                raise Unsupported(f"AggregatedIndexParameter agg_func {func!r}")
            refs = [self.index_param(c) for c in p.parameters]
            ref = self.emit(OP_AGG, [AGG[func], len(refs)] + refs)
        else:                                           # value() of the remaining IndexParameters is float(index)
            ref = self.param(p)
        self.memo[key] = ref
        return ref

    def prev_flow(self, node, initial=0.0, steps=1, deficit=False) -> int:
        """ref of the node's flow (or deficit) ``steps`` timesteps ago (``AbstractNode._prev_flow`` for one step; ``initial`` during
        the first ``steps`` timesteps): op LAG on a flow / deficit recorder of the node, which compile_program finds or appends
        once the recorders are known"""
        key = ("lag", id(node), float(initial), int(steps), bool(deficit))
        if key in self.memo:
            return self.memo[key]
        if not hasattr(self, "lag_nodes"):
            self.lag_nodes = []
        self.lag_nodes.append((node, bool(deficit)))
        ref = self.emit(OP_LAG, [-len(self.lag_nodes), self.const(initial), int(steps)])   # recorder index patched in compile_program
        self.memo[key] = ref
        return ref

    def threshold(self, p, v0, v1) -> int:
        """value1 if predicate(X, threshold) else value0 (AbstractThresholdParameter.index, _thresholds.pyx:78-112)"""
        from .bridge import _bridge

        name = type(p).__name__
        if getattr(p, "ratchet", False) and not getattr(self, "_in_ratchet", False):
            # once triggered it stays triggered (:95-112): index(t) = 1 if index(t-1) == 1 else predicate(t); yesterday's index is
            # the lagged sample of a hidden recorder of this very slot
            key = ("ratchet", id(p))
            if key not in self.memo:
                cell = {"slot": None}
                if not hasattr(self, "lag_nodes"):
                    self.lag_nodes = []
                self.lag_nodes.append((cell, "slot"))
                prev = self.emit(OP_LAG, [-len(self.lag_nodes), self.const(0.0), 1])
                self._in_ratchet = True
                try:
                    now = self.threshold(p, self.const(0.0), self.const(1.0))
                finally:
                    self._in_ratchet = False
                cell["slot"] = self.memo[key] = self.emit(OP_THRESHOLD, [0, prev, PRED["EQ"], self.const(1.0), now, self.const(1.0)])
            return self.emit(OP_THRESHOLD, [0, self.memo[key], PRED["EQ"], self.const(1.0), v0, v1])
        pred = _bridge.threshold_predicate(p)
        thr = self.value_ref(p.threshold)
        if name == "ParameterThresholdParameter":
            return self.emit(OP_THRESHOLD, [0, self.param(p.param), pred, thr, v0, v1])
        if name == "StorageThresholdParameter":
            return self.emit(OP_THRESHOLD, [1, self.storage_of(p.storage), pred, thr, v0, v1])
        if name == "NodeThresholdParameter":            # _thresholds.pyx:159-188: yesterday's flow; index 0 on the first timestep
            inner = self.emit(OP_THRESHOLD, [0, self.prev_flow(p.node), pred, thr, v0, v1])
            first = self.table(np.array([[1.0 if i == 0 else 0.0] for i in range(self.T)]), None)
            return self.emit(OP_THRESHOLD, [0, first, PRED["EQ"], self.const(1.0), inner, v0])
        if name == "RecorderThresholdParameter":        # _thresholds.pyx:320-360: yesterday's row of an array recorder; initial_value on day one
            if not hasattr(self, "lag_nodes"):
                self.lag_nodes = []
            self.lag_nodes.append((p.recorder, "recorder"))
            x = self.emit(OP_LAG, [-len(self.lag_nodes), self.const(0.0), 1])
            inner = self.emit(OP_THRESHOLD, [0, x, pred, thr, v0, v1])
            first = self.table(np.array([[1.0 if i == 0 else 0.0] for i in range(self.T)]), None)
            return self.emit(OP_THRESHOLD, [0, first, PRED["EQ"], self.const(1.0), inner, v1 if int(p.initial_value) else v0])
        if name == "CurrentYearThresholdParameter":
            x = self.table(np.array([[float(ts.year)] for ts in self.timesteps]), None)
        else:                                           # CurrentOrdinalDayThresholdParameter
            x = self.table(np.array([[float(ts.datetime.toordinal())] for ts in self.timesteps]), None)
        return self.emit(OP_THRESHOLD, [0, x, pred, thr, v0, v1])

    def storage_of(self, node) -> int:
        k = self.storage_index.get(id(node))
        if k is None and type(node).__name__ == "AggregatedStorage":
            k = self.aggregated_storage(node)
        if k is None:
            raise Unsupported(f"control curve on {type(node).__name__} {getattr(node, 'name', '')!r}")
        return k

    # -- aggregated storages ---------------------------------------------------------------------------
    def aggregated_storage(self, node) -> int:
        """An AggregatedStorage (pywr/_core.pyx:1363-1427) is not part of the LP: its volume follows the summed net inflow of
        its member storages, its proportion divides by their summed max_volume.  On the device it is one more storage
        without an LP row — kind "virtual" (mass balance from a linear form of the columns), always active — appended
        after the LP's own storages; control-curve parameters and storage recorders address it like any other."""
        k = self.storage_index.get(id(node))
        if k is not None:
            return k
        members = list(node.storage_nodes)
        for st in members:
            if self.storage_index.get(id(st)) is None or type(st).__name__ not in ("Storage", "Reservoir"):
                raise Unsupported(f"AggregatedStorage {node.name!r}: member {getattr(st, 'name', st)!r} is not a device storage")
        k = self.s.n_storage + len(self.extra_storages)
        self.storage_index[id(node)] = k
        self.extra_storages.append(node)
        return k

    def summed(self, refs) -> int:
        """ref of the sum of `refs` in member order (the order AggregatedStorage.after adds them in)"""
        if all(r < 0 for r in refs):
            total = 0.0
            for r in refs:
                total += self.consts[-r - 1]
            return self.const(total)
        return self.emit(OP_AGG, [AGG["sum"], len(refs)] + list(refs))

    # -- calendar-driven virtual storages ----------------------------------------------------------
    def schedule(self, node):
        """(reset[T], active[T]) of an Annual / Seasonal / MonthlyVirtualStorage, obtained by running the
        reference's OWN ``before(ts)`` over the calendar of the run (it depends on dates and the node's
        internal counters only, never on flows or volumes): reset[t] is 0, 1 (volume := max_volume) or
        2 (volume := initial volume), active[t] the ``active`` flag the LP row sees in timestep t."""
        from pywr.parameters import Parameter
        from .bridge import _bridge

        key = ("schedule", id(node))
        if key in self.memo:
            return self.memo[key]
        if isinstance(node.max_volume, Parameter):
            raise Unsupported(f"{type(node).__name__} with a max_volume parameter")
        sentinel = np.full(self.S, -1.2345e300)
        reset = np.zeros(self.T)
        active = np.zeros(self.T)
        node.reset()
        for t, ts in enumerate(self.timesteps):
            _bridge.set_storage_state(node, sentinel, sentinel)
            node.before(ts)
            if np.asarray(node.volume)[0] != sentinel[0]:
                reset[t] = 2.0 if getattr(node, "reset_to_initial_volume", False) else 1.0
            active[t] = 1.0 if node.active else 0.0
        node.reset()
        self.memo[key] = (reset, active)
        return self.memo[key]

    # -- bindings of the LP structure ------------------------------------------------------------
    def binding_ref(self, b) -> int:
        from pywr._core import StorageInput
        from pywr.parameters import Parameter

        kind, obj = b.kind, b.obj
        if b.loop:
            raise Unsupported(f"{type(obj).__name__}.get_{kind} is overridden in Python")
        if kind == "volume":
            return REF_STATE
        if kind == "active":
            if type(obj).__name__ in SCHEDULED_STORAGES:
                return self.table(self.schedule(obj)[1].reshape(-1, 1), None)
            if type(obj).__name__ not in ("VirtualStorage", "RollingVirtualStorage"):
                raise Unsupported(f"{type(obj).__name__}: activation / reset schedule is host-side")
            return self.const(1.0)
        if kind == "factor_norm":
            f = obj.factors
            return self.emit(OP_DIV, [self.param(f[0]), self.param(f[b.index])])
        if kind == "cost" and type(obj).__name__ in ("StorageInput", "StorageOutput"):
            value = obj.parent.cost
            if isinstance(value, Parameter):
                ref = self.param(value)
                return self.emit(OP_NEG, [ref]) if isinstance(obj, StorageInput) else ref
            return self.const(-value if isinstance(obj, StorageInput) else value)
        value = getattr(obj, kind)
        if kind == "min_flow" and value is None:
            value = 0.0
        return self.value_ref(value)


def _flow_form(structure: LPStructure, node, node_id, depth=0):
    """Linear form (col -> weight) of a node's flow after ``commit`` and ``after`` — the value the
    recorders read (``AbstractNode._flow``).  Plain nodes: the solver's own sum (cython_glpk.pyx:
    1078-1088).  Storage: outputs minus inputs (``_core.pyx:942-979``).  AggregatedNode: weighted
    member sum (``_core.pyx:762-776``).  Compound nodes copy a sub-node (``nodes.py:880-887``...)."""
    from pywr._core import AggregatedNode, AggregatedStorage, Storage, VirtualStorage

    s = structure
    form: Dict[int, float] = {}

    def add(other, w):
        for c, v in other.items():
            form[c] = form.get(c, 0.0) + w * v

    if depth > 8:
        raise Unsupported("node nesting too deep")
    if isinstance(node, AggregatedStorage):   # after(): sum of the members' flows (pywr/_core.pyx:1409-1413)
        for st in node.storage_nodes:
            add(_flow_form(s, st, node_id, depth + 1), 1.0)
        return form
    if isinstance(node, VirtualStorage):
        for n, f in zip(node.nodes, node.factors):
            add(_flow_form(s, n, node_id, depth + 1), -float(f))
        return form
    if isinstance(node, Storage):
        for n in node.outputs:
            add(_flow_form(s, n, node_id, depth + 1), 1.0)
        for n in node.inputs:
            add(_flow_form(s, n, node_id, depth + 1), -1.0)
        return form
    if isinstance(node, AggregatedNode):
        weights = node.flow_weights
        if weights is None:
            weights = [1.0] * len(node.nodes)
        for n, w in zip(node.nodes, weights):
            add(_flow_form(s, n, node_id, depth + 1), float(w))
        return form
    cname = type(node).__name__
    if cname in ("PiecewiseLink", "RiverGauge", "MultiSplitLink", "RiverSplit", "RiverSplitWithGauge"):
        # after(): commit_all(sub-output flow)   (pywr/nodes.py:880-887)
        return _flow_form(s, node.output, node_id, depth + 1)
    if cname == "LossLink":
        return _flow_form(s, node.net, node_id, depth + 1)  # pywr/nodes.py:1361-1364
    if cname == "BreakLink":
        return _flow_form(s, node.link, node_id, depth + 1)  # pywr/nodes.py:1168-1171
    i = node_id.get(id(node))
    if i is None:
        raise Unsupported(f"node {getattr(node, 'name', node)!r} is not part of the LP")
    for k in range(s.flow_indptr[i], s.flow_indptr[i + 1]):
        c = int(s.flow_col[k])
        form[c] = form.get(c, 0.0) + float(s.flow_w[k])
    if not form and cname not in ("Input", "Output", "Link", "Catchment", "River", "StorageInput", "StorageOutput"):
        raise Unsupported(f"flow of {cname} {getattr(node, 'name', '')!r} is computed on the host")
    return form


def _remap(refs, mapping):
    return np.array([mapping.get(int(r), int(r)) if r >= 0 else int(r) for r in refs], np.int32)


def compile_program(model, formulation="route", recorders=None, allow_host_recorders=False, population_axis=None):
    """Returns ``(structure, program)`` for a model that has been ``setup()`` and ``reset()``.

    ``structure`` is the LP structure with every input plane replaced by an op output or a constant
    and the storage volumes device-resident; ``program`` holds ops, tables, storage linear forms,
    initial state and recorders.  Raises :class:`Unsupported` when the model needs the host.

    ``allow_host_recorders``: recorders the device cannot evaluate (event / parameter / custom Python
    recorders ...) do not raise; they are returned in ``program.host_recorders`` and the caller keeps
    them alive with the host mirror protocol (``device_run.run_on_device``)."""
    from pywr.parameters import Parameter
    from pywr.recorders import Recorder

    from .bridge import _bridge

    for comp in model.components:
        # a bare Component (neither Parameter nor Recorder) has before() / after() hooks the reference calls every timestep
        # (pywr/_component.pyx; tests/test_components.py): only the per-timestep run honours them
        if not isinstance(comp, (Parameter, Recorder)):
            raise Unsupported(f"component {getattr(comp, 'name', comp)!r} of type {type(comp).__name__} runs on the host every timestep")
    base = build_structure(model, formulation)
    C = _Compiler(model, base)
    C.population_axis = population_axis
    S, T = C.S, C.T
    node_id = {id(n): i for i, n in enumerate(base.all_nodes)}

    # ---- LP inputs --------------------------------------------------------------------------
    mapping = {}
    for b in base.bindings:
        if b.slot >= 0:
            mapping[b.slot] = C.binding_ref(b)
    max_flow_ref = {}
    for b in base.bindings:
        if b.kind == "max_flow":
            max_flow_ref[id(b.obj)] = mapping[b.slot] if b.slot >= 0 else b.ref

    # ---- storages -------------------------------------------------------------------------------
    vol0 = np.zeros((max(base.n_storage, 1), S))
    pc0 = np.zeros((max(base.n_storage, 1), S))
    sf_indptr, sf_col, sf_w = [0], [], []
    reset_ops = []
    roll_len = np.zeros(max(base.n_storage, 1), np.int32)
    roll_init = np.zeros(max(base.n_storage, 1), np.float64)
    for k, st in enumerate(base.storages):
        cname = type(st).__name__
        # KeatingAquifer (pywr/domains/groundwater.py): a Storage with several inputs whose bounds are parameters; no state of its own
        if cname not in ("Storage", "Reservoir", "VirtualStorage", "RollingVirtualStorage", "KeatingAquifer") + SCHEDULED_STORAGES:
            raise Unsupported(f"storage type {cname} keeps extra state on the host")
        vol0[k] = np.asarray(st._volume)
        pc0[k] = np.asarray(st._current_pc)
        if cname == "RollingVirtualStorage":
            roll_len[k] = int(st.timesteps) - 1
            roll_init[k] = float(_bridge.rolling_initial_utilisation(st))
        if cname in SCHEDULED_STORAGES:
            # node.before() of the reference resets the volume on calendar dates BEFORE the parameters of that
            # timestep are evaluated: one op at the head of the op list, fed by the precomputed schedule
            reset, _ = C.schedule(st)
            if reset.any():
                flag = C.table(reset.reshape(-1, 1), None)
                init_v = C.table(vol0[k].reshape(1, -1).copy(), COL_SCENARIO) if np.ptp(vol0[k]) > 0 else C.const(vol0[k][0])
                init_p = C.table(pc0[k].reshape(1, -1).copy(), COL_SCENARIO) if np.ptp(pc0[k]) > 0 else C.const(pc0[k][0])
                reset_ops.append((OP_STORAGE_RESET, C.new_slot(), [k, flag, init_v, init_p]))
        form = _flow_form(base, st, node_id)
        for c in sorted(form):
            sf_col.append(c)
            sf_w.append(form[c])
        sf_indptr.append(len(sf_col))

    # ---- recorders ------------------------------------------------------------------------------
    if recorders is None:
        recorders = list(model.recorders)
    rec_kind, rec_src, rec_ref, rec_factor, rec_indptr, rec_col, rec_w, rec_objs = [], [], [], [], [0], [], [], []
    rec_row = []
    flow_row = {name[len("flow:"):]: i for i, name in enumerate(base.row_names) if name.startswith("flow:")}
    node_kinds = {
        "NumpyArrayNodeRecorder": REC_NODE_FLOW, "NumpyArrayNodeDeficitRecorder": REC_NODE_DEFICIT,
        "NumpyArrayNodeSuppliedRatioRecorder": REC_NODE_SUPPLIED,
        "NumpyArrayNodeCurtailmentRatioRecorder": REC_NODE_CURTAIL,
        "TotalDeficitNodeRecorder": REC_TOTAL_DEFICIT, "TotalFlowNodeRecorder": REC_TOTAL_FLOW,
        "MeanFlowNodeRecorder": REC_MEAN_FLOW, "DeficitFrequencyNodeRecorder": REC_DEFICIT_FREQ,
        # subclasses of NumpyArrayNodeRecorder that only add a finish() over the recorded array (:814-990)
        "FlowDurationCurveRecorder": REC_NODE_FLOW, "SeasonalFlowDurationCurveRecorder": REC_NODE_FLOW,
        "FlowDurationCurveDeviationRecorder": REC_NODE_FLOW,
    }
    storage_array_recorders = ("NumpyArrayStorageRecorder", "StorageDurationCurveRecorder")   # the latter: finish() only (:1243-1302)
    needs_max_flow = {REC_NODE_DEFICIT, REC_NODE_SUPPLIED, REC_NODE_CURTAIL, REC_TOTAL_DEFICIT, REC_DEFICIT_FREQ}
    host_recorders = []
    for rec in recorders:
        cname = type(rec).__name__
        if cname == "AggregatedRecorder":
            continue  # reduces other recorders' values() on the host at read-out (a13)
        if cname in ("NumpyArrayParameterRecorder", "NumpyArrayIndexParameterRecorder"):
            # the value (index) a parameter took in each timestep (pywr/recorders/_recorders.pyx:1361-1424, 1480-1540): a sample of
            # its V slot, expressed as the deficit recorder of an EMPTY flow form (value - 0): no new recorder kind in the kernels
            try:
                if type(rec._param).__name__ == "DeficitParameter":
                    # the one parameter whose value changes in after(): a recorder sees TODAY's deficit of the node (the LP saw
                    # yesterday's; pywr/parameters/_parameters.pyx:1910-1957, tests/test_parameters.py::test_deficit_parameter)
                    dnode = rec._param.node
                    form = _flow_form(base, dnode, node_id)
                    if id(dnode) not in max_flow_ref:
                        raise Unsupported("DeficitParameter on a node without an LP max_flow")
                    rec_kind.append(REC_NODE_DEFICIT); rec_src.append(node_id.get(id(dnode), -1)); rec_ref.append(max_flow_ref[id(dnode)])
                    rec_factor.append(1.0)
                    rec_row.append(flow_row.get(getattr(dnode, "fully_qualified_name", None), -1) if id(dnode) in node_id else -1)
                    for c in sorted(form):
                        rec_col.append(c); rec_w.append(form[c])
                    rec_indptr.append(len(rec_col))
                    rec_objs.append(rec)
                    continue
                ref = C.param(rec._param) if cname == "NumpyArrayParameterRecorder" else C.index_param(rec._param)
            except Unsupported:
                if not allow_host_recorders:
                    raise
                host_recorders.append(rec)
                continue
            rec_kind.append(REC_NODE_DEFICIT); rec_src.append(-1); rec_ref.append(ref); rec_factor.append(1.0); rec_row.append(-1)
            rec_indptr.append(len(rec_col))
            rec_objs.append(rec)
            continue
        if cname in ("TotalParameterRecorder", "MeanParameterRecorder"):
            # sum over time of value * factor (* days with integrate=True; pywr/recorders/_recorders.pyx:2274-2354; the mean divides
            # in the reference's own finish()): the accumulating deficit recorder of an empty flow form adds V[ref] * days, so the
            # slot holds value * factor, divided by the timestep length where that must not count (exact for daily timesteps)
            try:
                ref = C.param(rec._param)
                if float(rec.factor) != 1.0:
                    ref = C.emit(OP_AGG, [AGG["product"], 2, ref, C.const(float(rec.factor))])
                if not (cname == "TotalParameterRecorder" and rec.integrate) and np.any(np.asarray(C.days, dtype=np.float64) != 1.0):
                    ref = C.emit(OP_DIV, [ref, C.table(np.asarray(C.days, dtype=np.float64).reshape(-1, 1), None)])
            except Unsupported:
                if not allow_host_recorders:
                    raise
                host_recorders.append(rec)
                continue
            rec_kind.append(REC_TOTAL_DEFICIT); rec_src.append(-1); rec_ref.append(ref); rec_factor.append(1.0); rec_row.append(-1)
            rec_indptr.append(len(rec_col))
            rec_objs.append(rec)
            continue
        if cname in ("NumpyArrayLevelRecorder", "NumpyArrayAreaRecorder"):
            # Storage.get_level / get_area (pywr/_core.pyx:1256-1280, recorders :1305-1358): the value of the storage's level /
            # area parameter (or constant) in each timestep - a parameter sample like the two above
            try:
                st_node = _bridge.recorder_node(rec)
                ref = C.value_ref(st_node.level if cname == "NumpyArrayLevelRecorder" else st_node.area)
            except Unsupported:
                if not allow_host_recorders:
                    raise
                host_recorders.append(rec)
                continue
            rec_kind.append(REC_NODE_DEFICIT); rec_src.append(-1); rec_ref.append(ref); rec_factor.append(1.0); rec_row.append(-1)
            rec_indptr.append(len(rec_col))
            rec_objs.append(rec)
            continue
        try:
            node = _bridge.recorder_node(rec)
        except Exception:
            node = None
        if allow_host_recorders and (node is None or (cname not in node_kinds and cname not in storage_array_recorders
                                                      and cname != "MinimumVolumeStorageRecorder")):
            host_recorders.append(rec)
            continue
        if allow_host_recorders:
            try:  # a supported class on a node the device cannot express goes to the host as well
                if cname in node_kinds:
                    _flow_form(base, node, node_id)
                    if node_kinds[cname] in needs_max_flow and id(node) not in max_flow_ref:
                        raise Unsupported("no LP max_flow")
                elif C.storage_index.get(id(node)) is None:
                    if type(node).__name__ != "AggregatedStorage":
                        raise Unsupported("not a device storage")
                    C.aggregated_storage(node)
            except Unsupported:
                host_recorders.append(rec)
                continue
        if cname in node_kinds:
            kind = node_kinds[cname]
            form = _flow_form(base, node, node_id)
            ref = 0
            if kind in needs_max_flow:
                if id(node) not in max_flow_ref:
                    raise Unsupported(f"{cname} on {type(node).__name__} without an LP max_flow")
                ref = max_flow_ref[id(node)]
            factor = float(getattr(rec, "factor", 1.0)) if kind in (REC_NODE_FLOW, REC_TOTAL_FLOW, REC_MEAN_FLOW) else 1.0
            rec_kind.append(kind); rec_src.append(node_id.get(id(node), -1)); rec_ref.append(ref)
            rec_factor.append(factor)
            # plain Input/Link/Output node: its flow is the activity of its own LP row
            rec_row.append(flow_row.get(getattr(node, "fully_qualified_name", None), -1) if id(node) in node_id else -1)
            for c in sorted(form):
                rec_col.append(c); rec_w.append(form[c])
        elif cname in storage_array_recorders or cname == "MinimumVolumeStorageRecorder":
            k = C.storage_index.get(id(node))
            if k is None and type(node).__name__ == "AggregatedStorage":
                k = C.aggregated_storage(node)
            if k is None:
                raise Unsupported(f"{cname} on {type(node).__name__}")
            if cname == "MinimumVolumeStorageRecorder":
                kind = REC_MIN_VOLUME
            else:
                kind = REC_STORAGE_PC if rec.proportional else REC_STORAGE_VOLUME
            rec_kind.append(kind); rec_src.append(k); rec_ref.append(0); rec_factor.append(1.0); rec_row.append(-1)
        else:
            raise Unsupported(f"recorder type {cname} is evaluated on the host")
        rec_indptr.append(len(rec_col))
        rec_objs.append(rec)

    # ---- lagged node flows (op LAG): the state is a flow recorder's [T, S] array; reuse the user's, else append a hidden one ----
    n_visible = len(rec_objs)
    rec_names_all = [str(getattr(r, "name", "")) for r in rec_objs]
    lag_rec = []
    for node, deficit in getattr(C, "lag_nodes", []):
        if deficit == "slot":      # a ratchet's own index of the day before: hidden sample of its slot (empty flow form)
            rec_kind.append(REC_NODE_DEFICIT); rec_src.append(-1); rec_ref.append(node["slot"]); rec_factor.append(1.0); rec_row.append(-1)
            rec_indptr.append(len(rec_col))
            rec_names_all.append(f"__ratchet__{len(rec_kind)}")
            lag_rec.append((len(rec_kind) - 1, node, deficit))
            continue
        if deficit == "recorder":  # RecorderThresholdParameter: the user's own array recorder, which has to be a device recorder
            found = next((r for r in range(n_visible) if rec_objs[r] is node and rec_kind[r] in ARRAY_KINDS), None)
            if found is None:
                raise Unsupported(f"RecorderThresholdParameter on recorder {getattr(node, 'name', node)!r}, which does not record an array on the device")
            lag_rec.append((found, node, deficit))
            continue
        want = REC_NODE_DEFICIT if deficit else REC_NODE_FLOW
        found = next((r for r in range(n_visible) if rec_kind[r] == want and rec_factor[r] == 1.0
                      and _bridge.recorder_node(rec_objs[r]) is node), None)
        if found is None:
            found = next((r for r, nd, df in lag_rec if nd is node and df == deficit), None)
        if found is None:
            form = _flow_form(base, node, node_id)
            ref = 0
            if deficit:
                if id(node) not in max_flow_ref:
                    raise Unsupported(f"DeficitParameter on {type(node).__name__} without an LP max_flow")
                ref = max_flow_ref[id(node)]
            rec_kind.append(want); rec_src.append(node_id.get(id(node), -1)); rec_ref.append(ref); rec_factor.append(1.0)
            rec_row.append(flow_row.get(getattr(node, "fully_qualified_name", None), -1) if id(node) in node_id else -1)
            for c in sorted(form):
                rec_col.append(c); rec_w.append(form[c])
            rec_indptr.append(len(rec_col))
            rec_names_all.append(f"__prev_{'deficit' if deficit else 'flow'}__{node.name}")
            found = len(rec_kind) - 1
        lag_rec.append((found, node, deficit))
    for kind, _, args in C.ops:
        if kind == OP_LAG and args[0] < 0:
            args[0] = lag_rec[-args[0] - 1][0]

    # ---- aggregated storages: device storages without an LP row (C.aggregated_storage) -------------
    st_kind_all = list(base.st_kind)
    st_minv_all = list(_remap(base.st_minv_ref, mapping))
    st_maxv_all = list(_remap(base.st_maxv_ref, mapping))
    st_act_all = list(_remap(base.st_active_ref, mapping))
    for node in C.extra_storages:
        members = [C.storage_index[id(st)] for st in node.storage_nodes]
        st_kind_all.append(ST_KIND_VIRTUAL)
        st_minv_all.append(C.summed([int(st_minv_all[k]) for k in members]))
        st_maxv_all.append(C.summed([int(st_maxv_all[k]) for k in members]))
        st_act_all.append(C.const(1.0))
        vol0 = np.vstack([vol0[: len(st_kind_all) - 1], np.asarray(node._volume)[None, :]])
        pc0 = np.vstack([pc0[: len(st_kind_all) - 1], np.asarray(node._current_pc)[None, :]])
        form = _flow_form(base, node, node_id)
        for c in sorted(form):
            sf_col.append(c)
            sf_w.append(form[c])
        sf_indptr.append(len(sf_col))
    n_storage_all = len(st_kind_all)
    if C.extra_storages:
        roll_len = np.concatenate([roll_len[: base.n_storage], np.zeros(len(C.extra_storages), np.int32)])
        roll_init = np.concatenate([roll_init[: base.n_storage], np.zeros(len(C.extra_storages))])

    # ---- assemble --------------------------------------------------------------------------------
    consts = np.asarray(C.consts, np.float64)
    st_vol = np.full(n_storage_all, REF_STATE, np.int32)
    s = LPStructure(
        formulation=base.formulation, n_rows=base.n_rows, n_cols=base.n_cols, n_nodes=base.n_nodes,
        cost_clip=base.cost_clip, a_indptr=base.a_indptr, a_col=base.a_col, a_val=base.a_val,
        row_kind=base.row_kind,
        row_lb_ref=np.where(base.row_kind == 0, _remap(base.row_lb_ref, mapping), base.row_lb_ref).astype(np.int32),
        row_ub_ref=np.where(base.row_kind == 0, _remap(base.row_ub_ref, mapping), base.row_ub_ref).astype(np.int32),
        st_kind=np.asarray(st_kind_all, np.int32), st_vol_ref=st_vol, st_minv_ref=np.asarray(st_minv_all, np.int32),
        st_maxv_ref=np.asarray(st_maxv_all, np.int32), st_active_ref=np.asarray(st_act_all, np.int32),
        cost_indptr=base.cost_indptr, cost_ref=_remap(base.cost_ref, mapping), cost_w=base.cost_w,
        dyn_a_nz=base.dyn_a_nz, dyn_a_ref=_remap(base.dyn_a_ref, mapping), dyn_a_scale=base.dyn_a_scale,
        flow_indptr=base.flow_indptr, flow_col=base.flow_col, flow_w=base.flow_w,
        consts=consts, n_slots=max(C.n_slots, 1),
        node_names=base.node_names, row_names=base.row_names, col_names=base.col_names,
        bindings=base.bindings, all_nodes=base.all_nodes, routes=base.routes, storages=list(base.storages) + list(C.extra_storages),
    )
    if reset_ops:
        # op order = Model.before(): node.before() (scheduled resets) first, then the parameters.  Table ops
        # depend on nothing and go to the very front (the CUDA kernel hoists them anyway).
        tables_first = [op for op in C.ops if op[0] == OP_TABLE]
        others = [op for op in C.ops if op[0] != OP_TABLE]
        C.ops = tables_first + [(kind, dst, [int(a) for a in args]) for kind, dst, args in reset_ops] + others
    op_arg_ptr, op_args = [0], []
    for _, _, args in C.ops:
        op_args.extend(args)
        op_arg_ptr.append(len(op_args))
    offsets, chunks, pos = [], [], 0
    for values, _ in C.tables:
        offsets.append(pos)
        chunks.append(values.reshape(-1))
        pos += values.size
    i32 = lambda x: np.asarray(x, dtype=np.int32)  # noqa: E731
    f64 = lambda x: np.asarray(x, dtype=np.float64)  # noqa: E731
    prog = Program(
        n_steps=T, n_axes=C.n_axes,
        ts_days=f64(C.days),
        op_kind=i32([k for k, _, _ in C.ops]), op_dst=i32([d for _, d, _ in C.ops]),
        op_arg_ptr=i32(op_arg_ptr), op_args=i32(op_args),
        table_rows=i32([v.shape[0] for v, _ in C.tables]), table_cols=i32([v.shape[1] for v, _ in C.tables]),
        table_axis=i32([a for _, a in C.tables]), table_offset=np.asarray(offsets, np.int64),
        table_data=np.concatenate(chunks) if chunks else np.zeros(0),
        scen_index=C.scen_index.reshape(-1),
        stflow_indptr=i32(sf_indptr), stflow_col=i32(sf_col), stflow_w=f64(sf_w),
        initial_volume=vol0.reshape(-1), initial_pc=pc0.reshape(-1),
        rec_kind=i32(rec_kind), rec_src=i32(rec_src), rec_ref=i32(rec_ref), rec_factor=f64(rec_factor), rec_row=i32(rec_row),
        rec_indptr=i32(rec_indptr), rec_col=i32(rec_col), rec_w=f64(rec_w), recorders=rec_objs,
        rec_names=rec_names_all,
    )
    prog.host_recorders = host_recorders
    prog.variable_tables = dict(C.variable_tables)
    prog.months = np.asarray(C.months, np.int32)
    if roll_len.any():
        prog.st_roll_len, prog.st_roll_init = roll_len, roll_init
    return s, prog
