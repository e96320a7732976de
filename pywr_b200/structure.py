"""LP structure compiler: Pywr model graph -> one sparse constraint matrix shared by all scenarios.

Replaces ``CythonGLPKSolver.setup`` (reference ``pywr/solvers/cython_glpk.pyx:388-755``, route
formulation) and ``CythonGLPKEdgeSolver.setup`` (``:1148-1567``, edge formulation).  Instead of
filling a GLPK problem object it produces plain numpy arrays (:class:`LPStructure`) that are handed
to the CUDA engine through the C-ABI in ``include/b200lp.h`` (``b200lp_structure``), together with
the list of *bindings* that tell the host solver which ``Node`` / ``Storage`` attribute feeds which
per-scenario input plane ("slot") every timestep.

Row order, column order and node ids are those of the reference, because the per-scenario basis
arrays (``BasisManager``, ``cython_glpk.pyx:197-232``) are indexed by them.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Any, List, Optional

import numpy as np

ABI_VERSION = 8
ROW_FLOW = 0
ROW_STORAGE = 1
ST_KIND_STORAGE = 0
ST_KIND_VIRTUAL = 1
REF_STATE = -(2**31)


class b200lp_structure(ctypes.Structure):
    """ctypes mirror of ``struct b200lp_structure`` (include/b200lp.h)."""

    _i32p = ctypes.POINTER(ctypes.c_int32)
    _f64p = ctypes.POINTER(ctypes.c_double)
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("n_rows", ctypes.c_int32),
        ("n_cols", ctypes.c_int32),
        ("n_nodes", ctypes.c_int32),
        ("n_slots", ctypes.c_int32),
        ("n_consts", ctypes.c_int32),
        ("n_storage", ctypes.c_int32),
        ("n_dyn_a", ctypes.c_int32),
        ("cost_clip", ctypes.c_int32),
        ("reserved0", ctypes.c_int32),
        ("a_indptr", _i32p),
        ("a_col", _i32p),
        ("a_val", _f64p),
        ("row_kind", _i32p),
        ("row_lb_ref", _i32p),
        ("row_ub_ref", _i32p),
        ("st_kind", _i32p),
        ("st_vol_ref", _i32p),
        ("st_minv_ref", _i32p),
        ("st_maxv_ref", _i32p),
        ("st_active_ref", _i32p),
        ("cost_indptr", _i32p),
        ("cost_ref", _i32p),
        ("cost_w", _f64p),
        ("dyn_a_nz", _i32p),
        ("dyn_a_ref", _i32p),
        ("dyn_a_scale", _f64p),
        ("flow_indptr", _i32p),
        ("flow_col", _i32p),
        ("flow_w", _f64p),
        ("consts", _f64p),
    ]


_I32_FIELDS = (
    "a_indptr a_col row_kind row_lb_ref row_ub_ref st_kind st_vol_ref st_minv_ref st_maxv_ref "
    "st_active_ref cost_indptr cost_ref dyn_a_nz dyn_a_ref flow_indptr flow_col"
).split()
_F64_FIELDS = "a_val cost_w dyn_a_scale flow_w consts".split()


@dataclass(eq=False)
class Binding:
    """One per-scenario quantity the solver pulls from the model every timestep.

    ``kind`` is one of ``min_flow max_flow cost min_volume max_volume volume active factor_norm``;
    ``obj`` is the Node / Storage / AggregatedNode it is read from; ``index`` selects the factor for
    ``factor_norm``.  The value goes either to slot ``slot`` (per-scenario plane) or, when the
    attribute is a literal constant, to ``consts[const]``.
    """

    kind: str
    obj: Any
    slot: int = -1
    const: int = -1
    index: int = 0
    loop: bool = False  # a Python subclass overrides the scalar getter: call it per scenario
    baked: float = float("nan")  # factor_check: the normalised factor written into the matrix at setup

    @property
    def ref(self) -> int:
        return self.slot if self.slot >= 0 else -self.const - 1


@dataclass
class LPStructure:
    formulation: str
    n_rows: int
    n_cols: int
    n_nodes: int
    cost_clip: int
    a_indptr: np.ndarray
    a_col: np.ndarray
    a_val: np.ndarray
    row_kind: np.ndarray
    row_lb_ref: np.ndarray
    row_ub_ref: np.ndarray
    st_kind: np.ndarray
    st_vol_ref: np.ndarray
    st_minv_ref: np.ndarray
    st_maxv_ref: np.ndarray
    st_active_ref: np.ndarray
    cost_indptr: np.ndarray
    cost_ref: np.ndarray
    cost_w: np.ndarray
    dyn_a_nz: np.ndarray
    dyn_a_ref: np.ndarray
    dyn_a_scale: np.ndarray
    flow_indptr: np.ndarray
    flow_col: np.ndarray
    flow_w: np.ndarray
    consts: np.ndarray
    n_slots: int
    node_names: List[str] = field(default_factory=list)
    row_names: List[str] = field(default_factory=list)
    col_names: List[str] = field(default_factory=list)
    # host-side objects, not serialised
    bindings: List[Binding] = field(default_factory=list, repr=False)
    all_nodes: List[Any] = field(default_factory=list, repr=False)
    routes: Optional[list] = field(default=None, repr=False)
    storages: List[Any] = field(default_factory=list, repr=False)

    @property
    def n_storage(self) -> int:
        return int(self.st_kind.shape[0])

    @property
    def nnz(self) -> int:
        return int(self.a_val.shape[0])

    def dense(self) -> np.ndarray:
        A = np.zeros((self.n_rows, self.n_cols))
        for i in range(self.n_rows):
            for k in range(self.a_indptr[i], self.a_indptr[i + 1]):
                A[i, self.a_col[k]] += self.a_val[k]
        return A

    # ---- C-ABI marshalling -------------------------------------------------------------
    def to_c(self) -> b200lp_structure:
        cs = b200lp_structure()
        cs.abi_version = ABI_VERSION
        cs.n_rows, cs.n_cols, cs.n_nodes = self.n_rows, self.n_cols, self.n_nodes
        cs.n_slots, cs.n_consts = self.n_slots, int(self.consts.shape[0])
        cs.n_storage, cs.n_dyn_a = self.n_storage, int(self.dyn_a_nz.shape[0])
        cs.cost_clip = self.cost_clip
        keep = []
        for name in _I32_FIELDS:
            arr = np.ascontiguousarray(getattr(self, name), dtype=np.int32)
            if arr.size == 0:
                arr = np.zeros(1, np.int32)
            keep.append(arr)
            setattr(cs, name, arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        for name in _F64_FIELDS:
            arr = np.ascontiguousarray(getattr(self, name), dtype=np.float64)
            if arr.size == 0:
                arr = np.zeros(1, np.float64)
            keep.append(arr)
            setattr(cs, name, arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        cs._keepalive = keep  # the arrays must outlive the struct
        return cs

    # ---- serialisation (golden fixtures travel to the GPU box without the model) ------
    def save(self, path) -> None:
        data = {name: np.asarray(getattr(self, name)) for name in _I32_FIELDS + _F64_FIELDS}
        data["meta"] = np.array([self.n_rows, self.n_cols, self.n_nodes, self.cost_clip, self.n_slots], np.int64)
        data["formulation"] = np.array(self.formulation)
        data["node_names"] = np.array(self.node_names)
        data["row_names"] = np.array(self.row_names)
        data["col_names"] = np.array(self.col_names)
        np.savez_compressed(path, **data)

    @classmethod
    def load(cls, path) -> "LPStructure":
        z = np.load(path, allow_pickle=False)
        n_rows, n_cols, n_nodes, cost_clip, n_slots = (int(v) for v in z["meta"])
        kw = {name: z[name].astype(np.int32) for name in _I32_FIELDS}
        kw.update({name: z[name].astype(np.float64) for name in _F64_FIELDS})
        return cls(
            formulation=str(z["formulation"]),
            n_rows=n_rows, n_cols=n_cols, n_nodes=n_nodes, cost_clip=cost_clip, n_slots=n_slots,
            node_names=[str(s) for s in z["node_names"]],
            row_names=[str(s) for s in z["row_names"]],
            col_names=[str(s) for s in z["col_names"]],
            **kw,
        )


class _Builder:
    """Accumulates rows, refs and bindings while walking the model."""

    def __init__(self):
        self.rows = []  # list of dict col -> val
        self.row_kind, self.row_lb_ref, self.row_ub_ref, self.row_names = [], [], [], []
        self.consts: List[float] = []
        self.bindings: List[Binding] = []
        self.n_slots = 0
        self.dyn = []  # (row, col, ref, scale)

    def const_ref(self, value: float) -> int:
        self.consts.append(float(value))
        return -len(self.consts)

    def bind(self, kind, obj, value, index=0, force_slot=False) -> Binding:
        """Constant attribute -> const (re-checked by the solver each step); Parameter -> slot."""
        from pywr.parameters import Parameter

        b = Binding(kind=kind, obj=obj, index=index, loop=_overrides_getter(obj, kind))
        if force_slot or isinstance(value, Parameter) or b.loop:
            b.slot = self.n_slots
            self.n_slots += 1
        else:
            self.consts.append(float(value))
            b.const = len(self.consts) - 1
        self.bindings.append(b)
        return b

    def add_row(self, cols: dict, kind: int, lb_ref: int, ub_ref: int, name: str) -> int:
        self.rows.append(dict(cols))
        self.row_kind.append(kind)
        self.row_lb_ref.append(lb_ref)
        self.row_ub_ref.append(ub_ref)
        self.row_names.append(name)
        return len(self.rows) - 1


_GETTERS = {"min_flow": "get_min_flow", "max_flow": "get_max_flow", "cost": "get_cost",
            "min_volume": "get_min_volume", "max_volume": "get_max_volume"}


def _overrides_getter(obj, kind) -> bool:
    """True when a Python subclass overrides the scalar getter (e.g. pywr/nodes.py:834-838): the
    vectorised ``get_all_*`` would then bypass the override, so the solver must call it per scenario."""
    name = _GETTERS.get(kind)
    if name is None:
        return False
    for klass in type(obj).__mro__:
        if name in klass.__dict__:
            mod = getattr(klass, "__module__", "")
            return not mod.startswith("pywr._core")
    return False


def _attr(obj, kind):
    return getattr(obj, kind)


def _cross_domain_routes(model, BaseOutput, BaseInput):
    """``model.find_all_routes(BaseOutput, BaseInput, max_length=2, domain_match="different")`` (cython_glpk.pyx:426,1192) without
    its |outputs| x |inputs| calls of ``networkx.all_simple_paths`` (pywr/_model.pyx:543-546; 918 k calls and 7 s on the 2,073-node
    network): a route of at most two nodes is a direct edge, so the result is every edge Output -> Input whose ends lie in
    different domains, in the same order (both ends sorted by name)."""
    nodes = sorted(model.graph.nodes(), key=lambda n: n.name)
    inputs = [n for n in nodes if isinstance(n, BaseInput)]
    routes = []
    for out in nodes:
        if not isinstance(out, BaseOutput):
            continue
        succ = set(model.graph.successors(out))
        if not succ:
            continue
        for inp in inputs:
            if inp in succ and inp is not out and inp.domain != out.domain:
                routes.append([out, inp])
    return routes


def build_structure(model, formulation: str = "route") -> LPStructure:
    """Build the shared LP for ``model``.  Raises ``ModelStructureError`` exactly where the
    reference does (``cython_glpk.pyx:411,453,455,529`` / ``:1174,1222,1303,1308,1315``)."""
    from pywr._core import (AggregatedNode, BaseInput, BaseLink, BaseOutput, Storage, VirtualStorage)
    from pywr.core import ModelStructureError

    if formulation not in ("route", "edge"):
        raise ValueError(f"unknown formulation {formulation!r}")
    edge = formulation == "edge"

    all_nodes = list(sorted(model.graph.nodes(), key=lambda n: n.fully_qualified_name))
    if not all_nodes:
        raise ModelStructureError("Model is empty")
    node_id = {node: i for i, node in enumerate(all_nodes)}
    N = len(all_nodes)

    if edge:
        edges = list(model.graph.edges())
        if not edges:
            raise ModelStructureError("Model is empty")
        routes = None
        n_cols = len(edges)
        in_edges = {node: [] for node in all_nodes}
        out_edges = {node: [] for node in all_nodes}
        for col, (a, b) in enumerate(edges):
            if a.domain != b.domain:
                continue  # cross-domain edges stay empty columns (cython_glpk.pyx:1235-1239)
            out_edges[a].append(col)
            in_edges[b].append(col)
    else:
        routes = model.find_all_routes(BaseInput, BaseOutput, valid=(BaseLink, BaseInput, BaseOutput))
        n_cols = len(routes)
    cross_domain_routes = _cross_domain_routes(model, BaseOutput, BaseInput)

    non_storages, link_nodes, storages, virtual_storages, aggregated, aggregated_with_factors = [], [], [], [], [], []
    for node in all_nodes:
        if isinstance(node, (BaseInput, BaseLink, BaseOutput)):
            non_storages.append(node)
            if isinstance(node, BaseLink):
                link_nodes.append(node)
        elif isinstance(node, VirtualStorage):
            virtual_storages.append(node)
        elif isinstance(node, Storage):
            storages.append(node)
        elif isinstance(node, AggregatedNode):
            if node.factors is not None:
                aggregated_with_factors.append(node)
            aggregated.append(node)

    if not edge and len(routes) == 0:
        raise ModelStructureError("Model has no valid routes")
    if len(non_storages) == 0:
        raise ModelStructureError("Model has no non-storage nodes")

    def node_cols(node) -> list:
        """Columns that carry the flow of ``node`` (the set its flow constraint is written on)."""
        if edge:
            if isinstance(node, BaseOutput):
                return list(in_edges[node])
            return list(out_edges[node])
        if isinstance(node, BaseInput):
            return [c for c, r in enumerate(routes) if r[0] is node]
        if isinstance(node, BaseOutput):
            return [c for c, r in enumerate(routes) if r[-1] is node]
        return [c for c, r in enumerate(routes) if node in r]

    def member_cols(node) -> list:
        """Columns through ``node`` as used for aggregated / virtual-storage rows."""
        if edge:
            return node_cols(node)
        return [c for c, r in enumerate(routes) if node in r]

    B = _Builder()

    # cross-domain lookup (cython_glpk.pyx:475-487 / :1242-1254)
    cross_domain_cols = {}
    for output, inp in cross_domain_routes:
        conv = inp.get_conversion_factor()
        if edge:
            input_cols = [(c, conv) for c in out_edges[inp]]
        else:
            input_cols = [(c, conv) for c, r in enumerate(routes) if r[0] is inp]
        cross_domain_cols.setdefault(output, []).extend(input_cols)

    # --- non-storage rows ---------------------------------------------------------------
    cd_rows = []
    for node in non_storages:
        if edge:
            if isinstance(node, BaseInput) and in_edges[node]:
                raise ModelStructureError(f'Input node "{node.name}" should not have any upstream connections.')
            if isinstance(node, BaseOutput) and out_edges[node]:
                raise ModelStructureError(f'Output node "{node.name}" should not have any downstream connections.')
        cols = node_cols(node)
        if len(cols) == 0:
            if edge:
                raise ModelStructureError(
                    f'{node.__class__.__name__} node "{node.name}" does not have any valid connections.')
            raise ModelStructureError(f'{node.__class__.__name__} node "{node.name}" is not part of any routes.')
        lb = B.bind("min_flow", node, _attr(node, "min_flow"))
        ub = B.bind("max_flow", node, _attr(node, "max_flow"))
        B.add_row({c: 1.0 for c in cols}, ROW_FLOW, lb.ref, ub.ref, f"flow:{node.fully_qualified_name}")
        if node in cross_domain_cols:
            row = {}
            for c in cols:
                row[c] = row.get(c, 0.0) - 1.0
            for c, v in cross_domain_cols[node]:
                row[c] = row.get(c, 0.0) + 1.0 / v
            cd_rows.append((row, f"xdomain:{node.fully_qualified_name}"))
    zero = B.const_ref(0.0)
    for row, name in cd_rows:
        B.add_row(row, ROW_FLOW, zero, zero, name)

    # --- link mass balance (edge form only, :1364-1388) -----------------------------------
    if edge:
        for node in link_nodes:
            row = {}
            for c in in_edges[node]:
                row[c] = row.get(c, 0.0) + 1.0
            for c in out_edges[node]:
                row[c] = row.get(c, 0.0) - 1.0
            B.add_row(row, ROW_FLOW, zero, zero, f"balance:{node.fully_qualified_name}")

    # --- storage rows (:578-601 / :1393-1409) and virtual storage rows (:604-635 / :1424-1454)
    st_kind, st_vol_ref, st_minv_ref, st_maxv_ref, st_active_ref, st_objs = [], [], [], [], [], []

    def add_storage(storage, row, kind):
        k = len(st_kind)
        vol = B.bind("volume", storage, None, force_slot=True)
        minv = B.bind("min_volume", storage, _attr(storage, "min_volume"))
        maxv = B.bind("max_volume", storage, _attr(storage, "max_volume"))
        if kind == ST_KIND_VIRTUAL:
            act = B.bind("active", storage, None, force_slot=True).ref
        else:
            act = B.const_ref(1.0)
        st_kind.append(kind); st_vol_ref.append(vol.ref); st_minv_ref.append(minv.ref)
        st_maxv_ref.append(maxv.ref); st_active_ref.append(act); st_objs.append(storage)
        B.add_row(row, ROW_STORAGE, k, 0, f"storage:{storage.fully_qualified_name}")

    for storage in storages:
        row = {}
        if edge:
            for output in storage.outputs:
                for c in in_edges[output]:
                    row[c] = row.get(c, 0.0) + 1.0
            for inp in storage.inputs:
                for c in out_edges[inp]:
                    row[c] = row.get(c, 0.0) - 1.0
        else:
            for c, r in enumerate(routes):
                if r[-1] in storage.outputs and r[0] not in storage.inputs:
                    row[c] = row.get(c, 0.0) + 1.0
            for c, r in enumerate(routes):
                if r[0] in storage.inputs and r[-1] not in storage.outputs:
                    row[c] = row.get(c, 0.0) - 1.0
        add_storage(storage, row, ST_KIND_STORAGE)

    for storage in virtual_storages:
        cols = {}
        factors = storage.factors
        if edge:
            for i, some_node in enumerate(storage.nodes):
                for c in node_cols(some_node):
                    cols[c] = cols.get(c, 0.0) + factors[i]
        else:
            vs_nodes = list(storage.nodes)
            for c, r in enumerate(routes):
                for some_node in r:
                    try:
                        i = vs_nodes.index(some_node)
                    except ValueError:
                        continue
                    cols[c] = cols.get(c, 0.0) + factors[i]
        add_storage(storage, {c: -f for c, f in cols.items()}, ST_KIND_VIRTUAL)

    # --- aggregated nodes with factors (:639-673 / :1458-1499) ---------------------------------
    for agg in aggregated_with_factors:
        nodes = agg.nodes
        first_cols = member_cols(nodes[0])
        constant = bool(agg.has_constant_factors)
        norm = np.asarray(agg.get_factors_norm(None)) if constant else None
        for i, node in enumerate(nodes[1:], start=1):
            cols_i = member_cols(node)
            row = {}
            for c in first_cols:
                row[c] = row.get(c, 0.0) + 1.0
            r_index = len(B.rows)
            if constant:
                for c in cols_i:
                    row[c] = row.get(c, 0.0) - float(norm[i])
                # the reference re-reads get_factors_norm every scenario and timestep (cython_glpk.pyx:663-673) unless
                # set_fixed_factors_once; the solver compares with `baked` and rebuilds when a constant factor was changed
                B.bindings.append(Binding(kind="factor_check", obj=agg, index=i, const=-1, baked=float(norm[i])))
            else:
                b = B.bind("factor_norm", agg, None, index=i, force_slot=True)
                for c in cols_i:
                    if c in row:
                        raise ModelStructureError(
                            f'AggregatedNode "{agg.name}": nodes "{nodes[0].name}" and "{node.name}" share a column; '
                            "dynamic factors cannot be applied")
                    row[c] = -1.0
                    B.dyn.append((r_index, c, b.ref, -1.0))
            B.add_row(row, ROW_FLOW, zero, zero, f"factor:{agg.fully_qualified_name}:{i}")

    # --- aggregated min/max (:676-706 / :1502-1537) ----------------------------------------------
    for agg in aggregated:
        nodes = agg.nodes
        weights = agg.flow_weights
        if weights is None:
            weights = [1.0] * len(nodes)
        matrix = {}
        for some_node, w in zip(nodes, weights):
            for c in member_cols(some_node):
                matrix[c] = float(w)
        lb = B.bind("min_flow", agg, _attr(agg, "min_flow"))
        ub = B.bind("max_flow", agg, _attr(agg, "max_flow"))
        B.add_row({c: matrix[c] for c in sorted(matrix)}, ROW_FLOW, lb.ref, ub.ref,
                  f"aggregated:{agg.fully_qualified_name}")

    # --- assemble CSR ----------------------------------------------------------------------------
    a_indptr, a_col, a_val = [0], [], []
    nz_index = {}
    for i, row in enumerate(B.rows):
        for c in sorted(row):
            nz_index[(i, c)] = len(a_col)
            a_col.append(c)
            a_val.append(row[c])
        a_indptr.append(len(a_col))
    dyn_a_nz = [nz_index[(r, c)] for r, c, _, _ in B.dyn]
    dyn_a_ref = [ref for _, _, ref, _ in B.dyn]
    dyn_a_scale = [sc for _, _, _, sc in B.dyn]

    # --- objective (:709-724 / :1283-1289,1701-1714) ---------------------------------------------
    cost_bind = {}

    def cost_ref_of(node):
        if node not in cost_bind:
            cost_bind[node] = B.bind("cost", node, _node_cost_attr(node))
        return cost_bind[node].ref

    cost_terms = [[] for _ in range(n_cols)]
    if edge:
        for node in all_nodes:
            if not (in_edges[node] or out_edges[node]):
                continue
            if not hasattr(node, "get_cost"):
                continue
            w = 0.5 if isinstance(node, BaseLink) else 1.0
            ref = cost_ref_of(node)
            for c in in_edges[node]:
                cost_terms[c].append((ref, w))
            for c in out_edges[node]:
                cost_terms[c].append((ref, w))
    else:
        for c, r in enumerate(routes):
            cost_terms[c].append((cost_ref_of(r[0]), 1.0))
            for some_node in r[1:-1]:
                if isinstance(some_node, BaseLink):
                    cost_terms[c].append((cost_ref_of(some_node), 1.0))
            cost_terms[c].append((cost_ref_of(r[-1]), 1.0))
    cost_indptr, cost_ref, cost_w = [0], [], []
    for terms in cost_terms:
        for ref, w in terms:
            cost_ref.append(ref)
            cost_w.append(w)
        cost_indptr.append(len(cost_ref))

    # --- result extraction (:1078-1088 / :1899-1911) ---------------------------------------------
    flow_terms = [[] for _ in range(N)]
    if edge:
        for c, (a, b) in enumerate(edges):
            for node in (a, b):
                w = 0.5 if isinstance(node, BaseLink) else 1.0
                flow_terms[node_id[node]].append((c, w))
    else:
        for c, r in enumerate(routes):
            last = len(r) - 1
            for pos, node in enumerate(r):
                if pos == 0 or pos == last or isinstance(node, BaseLink):
                    flow_terms[node_id[node]].append((c, 1.0))
    flow_indptr, flow_col, flow_w = [0], [], []
    for terms in flow_terms:
        for c, w in terms:
            flow_col.append(c)
            flow_w.append(w)
        flow_indptr.append(len(flow_col))

    if edge:
        col_names = [f"{a.fully_qualified_name}->{b.fully_qualified_name}" for a, b in edges]
    else:
        col_names = ["->".join(n.fully_qualified_name for n in r) for r in routes]

    i32 = lambda x: np.asarray(x, dtype=np.int32)  # noqa: E731
    f64 = lambda x: np.asarray(x, dtype=np.float64)  # noqa: E731
    s = LPStructure(
        formulation=formulation, n_rows=len(B.rows), n_cols=n_cols, n_nodes=N, cost_clip=0 if edge else 1,
        a_indptr=i32(a_indptr), a_col=i32(a_col), a_val=f64(a_val),
        row_kind=i32(B.row_kind), row_lb_ref=i32(B.row_lb_ref), row_ub_ref=i32(B.row_ub_ref),
        st_kind=i32(st_kind), st_vol_ref=i32(st_vol_ref), st_minv_ref=i32(st_minv_ref),
        st_maxv_ref=i32(st_maxv_ref), st_active_ref=i32(st_active_ref),
        cost_indptr=i32(cost_indptr), cost_ref=i32(cost_ref), cost_w=f64(cost_w),
        dyn_a_nz=i32(dyn_a_nz), dyn_a_ref=i32(dyn_a_ref), dyn_a_scale=f64(dyn_a_scale),
        flow_indptr=i32(flow_indptr), flow_col=i32(flow_col), flow_w=f64(flow_w),
        consts=f64(B.consts), n_slots=B.n_slots,
        node_names=[n.fully_qualified_name for n in all_nodes], row_names=B.row_names, col_names=col_names,
        bindings=B.bindings, all_nodes=all_nodes, routes=routes, storages=st_objs,
    )
    return s


def _node_cost_attr(node):
    """The attribute that decides constant-vs-parameter for a node's cost.  StorageInput /
    StorageOutput take the parent's cost with sign -1 / +1 (pywr/_core.pyx:959-961,993-995); the
    sign is applied by ``get_all_cost`` when the value is gathered, and by the constant below."""
    from pywr._core import StorageInput, StorageOutput
    from pywr.parameters import Parameter

    if isinstance(node, (StorageInput, StorageOutput)):
        parent_cost = node.parent.cost
        if isinstance(parent_cost, Parameter):
            return parent_cost
        return -parent_cost if isinstance(node, StorageInput) else parent_cost
    return node.cost
