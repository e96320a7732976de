"""ctypes binding of the C-ABI in ``include/b200lp.h`` (``libb200lp.so``, built from
``pywr_b200/csrc``).  This is the only way the Python host reaches the GPU: no torch, no CPU
fallback — if the library or a CUDA device is missing, construction raises."""

from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200LP_LIB") or os.path.join(_HERE, "libb200lp.so")  # B200LP_LIB: experiment builds
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


class EngineUnavailable(RuntimeError):
    pass


class b200lp_stats(ctypes.Structure):
    _fields_ = [
        ("steps", ctypes.c_int64), ("scenario_steps", ctypes.c_int64), ("pivots", ctypes.c_int64),
        ("refactors", ctypes.c_int64), ("kernel_launches", ctypes.c_int64),
        ("h2d_bytes", ctypes.c_double), ("d2h_bytes", ctypes.c_double), ("device_ms", ctypes.c_double),
        ("n_rows", ctypes.c_int32), ("n_cols", ctypes.c_int32), ("n_nonzero", ctypes.c_int32),
        ("kernel_variant", ctypes.c_int32),
        ("pdhg_iterations", ctypes.c_int64), ("reduced_rows", ctypes.c_int32), ("reduced_nonzero", ctypes.c_int32),
        ("run_kernel_jit", ctypes.c_int32), ("reserved1", ctypes.c_int32),
    ]


# every symbol include/b200lp.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("b200lp_last_error", ctypes.c_char_p, []),
    ("b200lp_abi_version", ctypes.c_int, []),
    ("b200lp_device_count", ctypes.c_int, []),
    ("b200lp_create", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    ("b200lp_create_ex", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p)]),
    ("b200lp_destroy", None, [ctypes.c_void_p]),
    ("b200lp_reset", ctypes.c_int, [ctypes.c_void_p, _dp]),
    ("b200lp_step", ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, _dp, _dp, _dp, _dp]),
    ("b200lp_step_device", ctypes.c_int, [ctypes.c_void_p, ctypes.c_double] + [ctypes.c_void_p] * 4),
    ("b200lp_set_consts", ctypes.c_int, [ctypes.c_void_p, _dp]),
    ("b200lp_status", ctypes.c_int, [ctypes.c_void_p, _ip, _ip]),
    ("b200lp_get_basis", ctypes.c_int, [ctypes.c_void_p, _ip, _ip]),
    ("b200lp_get_volume", ctypes.c_int, [ctypes.c_void_p, _dp]),
    ("b200lp_set_volume", ctypes.c_int, [ctypes.c_void_p, _dp]),
    ("b200lp_get_stats", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(b200lp_stats)]),
    ("b200lp_set_option", ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]),
    ("b200lp_stream", ctypes.c_void_p, [ctypes.c_void_p]),
    ("b200lp_sync", ctypes.c_int, [ctypes.c_void_p]),
    ("b200lp_host_alloc", ctypes.c_void_p, [ctypes.c_size_t]),
    ("b200lp_host_free", None, [ctypes.c_void_p]),
    ("b200lp_device_alloc", ctypes.c_void_p, [ctypes.c_size_t]),
    ("b200lp_device_free", None, [ctypes.c_void_p]),
    ("b200lp_memcpy_h2d", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    ("b200lp_memcpy_d2h", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    ("b200lp_load_program", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    ("b200lp_jit_probe", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_char_p, ctypes.c_size_t]),
    ("b200lp_program_reset", ctypes.c_int, [ctypes.c_void_p]),
    ("b200lp_run", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]),
    ("b200lp_run_pipelined", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    ("b200lp_program_status", ctypes.c_int, [ctypes.c_void_p, _ip, _ip, _ip, _ip]),
    ("b200lp_update_table", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, _dp]),
    ("b200lp_mark", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    ("b200lp_elapsed_ms", ctypes.c_int, [ctypes.c_void_p, _dp]),
    ("b200lp_fetch_recorder", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, _dp]),
    ("b200lp_fetch_recorder_strided", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64]),
    ("b200lp_fetch_recorder_agg", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, _dp, _dp, _dp]),
    ("b200lp_recorder_device_ptr", ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)]),
    ("b200lp_fetch_state", ctypes.c_int, [ctypes.c_void_p, _dp, _dp, _dp]),
]


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library and bind every declared symbol.  Raises if it has not been built
    (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise EngineUnavailable(
            f"{path} not found: build the CUDA extension first (__graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(L, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    if L.b200lp_abi_version() != 8:
        raise EngineUnavailable("libb200lp.so ABI version mismatch: rebuild")
    _lib = L
    return L


def _check(L, rc, what):
    if rc != 0:
        msg = L.b200lp_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise EngineUnavailable(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


def _ptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


class PinnedArray:
    """float64 array over cudaHostAlloc'ed memory (for cudaMemcpyAsync at full PCIe rate)."""

    def __init__(self, L, shape):
        self.L = L
        n = int(np.prod(shape))
        self.ptr = L.b200lp_host_alloc(max(n, 1) * 8)
        if not self.ptr:
            raise MemoryError("b200lp_host_alloc failed")
        buf = (ctypes.c_double * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
        self.array[...] = 0.0

    def free(self):
        if self.ptr:
            self.array = None
            self.L.b200lp_host_free(self.ptr)
            self.ptr = None


class CudaEngine:
    """One engine per (LP structure, scenario count, device)."""

    def __init__(self, structure, n_scenarios, device=None, for_program=False, **_):
        self.L = load_library()
        self.structure = structure
        self.S = int(n_scenarios)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) % max(self.L.b200lp_device_count(), 1)
        self.device = int(device)
        self._cs = structure.to_c()
        h = ctypes.c_void_p()
        # for_program: B200LP_CREATE_FOR_PROGRAM — mid-size LPs take the large-LP path, which runs device-resident programs
        _check(self.L, self.L.b200lp_create_ex(ctypes.addressof(self._cs), self.S, self.device, 1 if for_program else 0,
                                               ctypes.byref(h)), "b200lp_create")
        self.h = h
        self._pinned = []

    # -- buffers ------------------------------------------------------------------------------
    def alloc_planes(self, n):
        pa = PinnedArray(self.L, (int(n), self.S))
        self._pinned.append(pa)
        return pa.array

    # -- lifecycle ----------------------------------------------------------------------------
    def reset(self, vol0=None):
        _check(self.L, self.L.b200lp_reset(self.h, _ptr(vol0)), "b200lp_reset")

    def set_consts(self, consts):
        _check(self.L, self.L.b200lp_set_consts(self.h, _ptr(np.ascontiguousarray(consts, dtype=np.float64))),
               "b200lp_set_consts")

    def step(self, days, slots, node_flow=None, col_flow=None, objective=None):
        """Returns the number of failed scenarios (0 = all optimal) like the oracle engine."""
        rc = self.L.b200lp_step(self.h, float(days), _ptr(slots), _ptr(node_flow), _ptr(col_flow), _ptr(objective))
        if rc == -5:
            return 1
        _check(self.L, rc, "b200lp_step")
        return 0

    def step_device(self, days, d_slots, d_node_flow=None, d_col_flow=None, d_objective=None):
        _check(self.L, self.L.b200lp_step_device(self.h, float(days), d_slots, d_node_flow, d_col_flow, d_objective),
               "b200lp_step_device")

    def status(self):
        a, b = ctypes.c_int32(), ctypes.c_int32()
        _check(self.L, self.L.b200lp_status(self.h, ctypes.byref(a), ctypes.byref(b)), "b200lp_status")
        return a.value, b.value

    def get_basis(self):
        s = self.structure
        row = np.zeros((self.S, s.n_rows), np.int32)
        col = np.zeros((self.S, s.n_cols), np.int32)
        _check(self.L, self.L.b200lp_get_basis(self.h, row.ctypes.data_as(_ip), col.ctypes.data_as(_ip)),
               "b200lp_get_basis")
        return row, col

    def get_volume(self):
        out = np.zeros((self.structure.n_storage, self.S))
        _check(self.L, self.L.b200lp_get_volume(self.h, _ptr(out)), "b200lp_get_volume")
        return out

    def stats(self):
        st = b200lp_stats()
        _check(self.L, self.L.b200lp_get_stats(self.h, ctypes.byref(st)), "b200lp_get_stats")
        return {name: getattr(st, name) for name, _ in b200lp_stats._fields_}

    def counters(self):
        st = self.stats()
        return {"pivots": st["pivots"], "refactors": st["refactors"], "steps": st["steps"]}

    def sync(self):
        _check(self.L, self.L.b200lp_sync(self.h), "b200lp_sync")

    def set_option(self, name, value):
        _check(self.L, self.L.b200lp_set_option(self.h, name.encode(), float(value)), "b200lp_set_option")

    # -- device-resident run -----------------------------------------------------------------
    def load_program(self, prog):
        self._cp = prog.to_c()
        self._prog = prog
        _check(self.L, self.L.b200lp_load_program(self.h, ctypes.addressof(self._cp)), "b200lp_load_program")

    def program_reset(self):
        _check(self.L, self.L.b200lp_program_reset(self.h), "b200lp_program_reset")

    def run(self, t0, n):
        """Queues timesteps [t0, t0+n) on the engine's stream (asynchronous)."""
        _check(self.L, self.L.b200lp_run(self.h, int(t0), int(n)), "b200lp_run")
        return 0

    def run_pipelined(self, tables=None, outs=None, chunks=16):
        """Whole run with host buffers (reset, upload, solve, read-out overlapped in time slices).
        tables: {table index: array}; outs: list with one array (or None) per recorder."""
        nt, nr = self._prog.n_tables, self._prog.n_recorders
        tp = (ctypes.c_void_p * max(nt, 1))()
        for k, a in (tables or {}).items():
            assert a.dtype == np.float64 and a.flags.c_contiguous
            tp[k] = a.ctypes.data
        rp = (ctypes.c_void_p * max(nr, 1))()
        for r, a in enumerate(outs or []):
            if a is not None:
                assert a.dtype == np.float64 and a.flags.c_contiguous
                rp[r] = a.ctypes.data
        _check(self.L, self.L.b200lp_run_pipelined(self.h, int(chunks), tp if tables else None, rp if outs else None),
               "b200lp_run_pipelined")

    def program_status(self):
        v = [ctypes.c_int32() for _ in range(4)]
        _check(self.L, self.L.b200lp_program_status(self.h, *[ctypes.byref(x) for x in v]), "b200lp_program_status")
        return tuple(x.value for x in v)

    def fetch_recorder(self, r, shape, out=None):
        if out is None:
            out = np.zeros(shape, dtype=np.float64)
        _check(self.L, self.L.b200lp_fetch_recorder(self.h, int(r), _ptr(out)), "b200lp_fetch_recorder")
        return out

    def fetch_recorder_into(self, r, out, col0):
        """Queues the copy of recorder ``r`` into columns [col0, col0 + S) of the C-contiguous host array ``out``
        ([T, S_total] or [S_total]); asynchronous: ``sync()`` before reading."""
        assert out.dtype == np.float64 and out.flags.c_contiguous
        ld = out.shape[-1]
        _check(self.L, self.L.b200lp_fetch_recorder_strided(self.h, int(r), out.ctypes.data + 8 * int(col0), int(ld)),
               "b200lp_fetch_recorder_strided")

    def fetch_recorder_agg(self, r):
        """(sum, min, max) over the executed timesteps of array recorder ``r``, [S] each, accumulated on the device."""
        out = np.zeros((3, self.S), dtype=np.float64)
        _check(self.L, self.L.b200lp_fetch_recorder_agg(self.h, int(r), _ptr(out[0]), _ptr(out[1]), _ptr(out[2])),
               "b200lp_fetch_recorder_agg")
        return out[0], out[1], out[2]

    def update_table(self, table, data):
        """Asynchronous upload of one table (use a pinned array from alloc_pinned for full PCIe rate)."""
        _check(self.L, self.L.b200lp_update_table(self.h, int(table), _ptr(data)), "b200lp_update_table")

    def alloc_pinned(self, shape):
        pa = PinnedArray(self.L, tuple(int(v) for v in shape))
        self._pinned.append(pa)
        return pa.array

    def mark(self, which):
        _check(self.L, self.L.b200lp_mark(self.h, int(which)), "b200lp_mark")

    def elapsed_ms(self):
        ms = ctypes.c_double()
        _check(self.L, self.L.b200lp_elapsed_ms(self.h, ctypes.byref(ms)), "b200lp_elapsed_ms")
        return ms.value

    def recorder_device_ptr(self, r):
        n = ctypes.c_int64()
        p = self.L.b200lp_recorder_device_ptr(self.h, int(r), ctypes.byref(n))
        return p, n.value

    def fetch_state(self):
        s = self.structure
        x = np.zeros((s.n_cols, self.S))
        vol = np.zeros((max(s.n_storage, 1), self.S))
        pc = np.zeros((max(s.n_storage, 1), self.S))
        _check(self.L, self.L.b200lp_fetch_state(self.h, _ptr(x), _ptr(vol), _ptr(pc)), "b200lp_fetch_state")
        return x, vol, pc

    @property
    def stream(self):
        return self.L.b200lp_stream(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.b200lp_destroy(self.h)
            self.h = None
        for pa in getattr(self, "_pinned", []):
            pa.free()
        self._pinned = []

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def scenario_blocks(n_scenarios, n_shards):
    """Contiguous blocks of ``global_id`` (pywr/_core.pyx:210-240), sizes differing by at most one."""
    base, extra = divmod(int(n_scenarios), int(n_shards))
    out, lo = [], 0
    for k in range(n_shards):
        hi = lo + base + (1 if k < extra else 0)
        out.append((lo, hi))
        lo = hi
    return [b for b in out if b[1] > b[0]]


class ShardedEngine:
    """One scenario batch over several engines — one per GPU of the box — driven by ONE host process
    (the ``devices`` of the boundary).  Scenarios are independent (cython_glpk.pyx:827-828), so the
    batch is cut into contiguous ``global_id`` blocks; every call fans out to the shards (their work is queued on
    per-device streams and runs concurrently), read-outs land in the right column block of the reference's
    ``[T, ncomb]`` / ``[ncomb]`` layout.  No data-path collective exists or is needed.

    ``devices``: ``"all"``, a count, or a list of CUDA device indices (a device may appear more than once: several
    handles on one GPU); ``factory(structure, n, device)`` builds one shard engine (default: :class:`CudaEngine`)."""

    def __init__(self, structure, n_scenarios, devices="all", for_program=False, factory=None):
        self.structure, self.S = structure, int(n_scenarios)
        if factory is None:
            L = load_library()
            if devices == "all":
                devices = list(range(max(L.b200lp_device_count(), 1)))

            def factory(s, n, device):
                return CudaEngine(s, n, device=device, for_program=for_program)
        if isinstance(devices, int):
            devices = list(range(devices))
        if devices == "all":
            raise ValueError("devices='all' needs the CUDA engine (no factory given)")
        devices = list(devices)[: self.S]
        self.blocks = scenario_blocks(self.S, len(devices))
        self.shards = []
        try:
            for (lo, hi), d in zip(self.blocks, devices):
                self.shards.append(factory(structure, hi - lo, d))
        except Exception:
            self.close()
            raise
        self.devices = devices[: len(self.shards)]

    # -- device-resident run ---------------------------------------------------------------------
    def load_program(self, prog):
        self._prog = prog
        for (lo, hi), e in zip(self.blocks, self.shards):
            e.load_program(prog.shard(lo, hi, self.S))

    def program_reset(self):
        for e in self.shards:
            e.program_reset()

    def run(self, t0, n):
        for e in self.shards:     # asynchronous launches: the shards run concurrently
            e.run(t0, n)
        return 0

    def sync(self):
        for e in self.shards:
            e.sync()

    def program_status(self):
        n_failed, first = 0, (-1, 0, -1)
        for (lo, _), e in zip(self.blocks, self.shards):
            n, scen, code, ts = e.program_status()
            n_failed += n
            if n and first[0] < 0:
                first = (lo + scen, code, ts)
        return (n_failed,) + first

    def fetch_recorder(self, r, shape, out=None):
        if out is None:
            out = np.zeros(shape, dtype=np.float64)
        for (lo, _), e in zip(self.blocks, self.shards):
            if hasattr(e, "fetch_recorder_into"):
                e.fetch_recorder_into(r, out, lo)
            else:                                   # engines without the strided copy (test doubles, oracle)
                part = e.fetch_recorder(r, shape[:-1] + (e.S,))
                out[..., lo: lo + e.S] = part
        self.sync()
        return out

    def fetch_recorder_agg(self, r):
        out = np.zeros((3, self.S))
        for (lo, hi), e in zip(self.blocks, self.shards):
            a, b, c = e.fetch_recorder_agg(r)
            out[0, lo:hi], out[1, lo:hi], out[2, lo:hi] = a, b, c
        return out[0], out[1], out[2]

    def fetch_state(self):
        parts = [e.fetch_state() for e in self.shards]
        return tuple(np.concatenate([p[k] for p in parts], axis=1) for k in range(3))

    def update_table(self, table, data):
        from .program import COL_SCENARIO

        per_scenario = int(self._prog.table_axis[table]) == COL_SCENARIO
        for (lo, hi), e in zip(self.blocks, self.shards):
            e.update_table(table, np.ascontiguousarray(data[:, lo:hi]) if per_scenario else data)

    def stats(self):
        out = None
        for e in self.shards:
            st = e.stats()
            if out is None:
                out = dict(st)
            else:
                for k in ("scenario_steps", "pivots", "refactors", "kernel_launches", "h2d_bytes", "d2h_bytes", "pdhg_iterations"):
                    if k in out:
                        out[k] += st[k]
        return out

    def counters(self):
        c = [e.counters() for e in self.shards]
        return {"pivots": sum(x["pivots"] for x in c), "refactors": sum(x["refactors"] for x in c), "steps": c[0]["steps"]}

    def close(self):
        for e in getattr(self, "shards", []):
            e.close()
        self.shards = []
