"""ctypes binding of the CPU ORACLE (oracle/lp_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` legs and ``__graft_entry__.smoke()``
may import this module; the product package ``pywr_b200`` never does.  The class has the same
duck-typed interface as ``pywr_b200.engine.CudaEngine`` so that the same host solver logic
(``pywr_b200.solver.BatchedLPSolver``) can drive either.
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liblporacle.so")
_lib = None


def build(force=False):
    """Compile the oracle with gcc (seconds)."""
    src = os.path.join(_HERE, "lp_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "b200lp.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "build/liblporacle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_void_p, ctypes.c_int32]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_reset.argtypes = [ctypes.c_void_p, dp]
        L.orc_step.restype = ctypes.c_int
        L.orc_step.argtypes = [ctypes.c_void_p, ctypes.c_double, dp, dp, dp, dp]
        L.orc_status.argtypes = [ctypes.c_void_p, ip, ip]
        L.orc_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_int32]
        L.orc_set_consts.argtypes = [ctypes.c_void_p, dp]
        L.orc_get_basis.argtypes = [ctypes.c_void_p, ip, ip]
        L.orc_get_volume.argtypes = [ctypes.c_void_p, dp]
        L.orc_set_volume.argtypes = [ctypes.c_void_p, dp]
        L.orc_counters.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int64)] * 3
        L.orc_load_program.restype = ctypes.c_int
        L.orc_load_program.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_program_reset.argtypes = [ctypes.c_void_p]
        L.orc_update_table.restype = ctypes.c_int
        L.orc_update_table.argtypes = [ctypes.c_void_p, ctypes.c_int32, dp]
        L.orc_run.restype = ctypes.c_int
        L.orc_run.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]
        L.orc_program_status.argtypes = [ctypes.c_void_p] + [ip] * 4
        L.orc_fetch_recorder.restype = ctypes.c_int
        L.orc_fetch_recorder.argtypes = [ctypes.c_void_p, ctypes.c_int32, dp]
        L.orc_fetch_state.restype = ctypes.c_int
        L.orc_fetch_state.argtypes = [ctypes.c_void_p, dp, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class OracleEngine:
    def __init__(self, structure, n_scenarios, threads=None, **_):
        self.L = lib()
        self.structure = structure
        self.S = int(n_scenarios)
        self._cs = structure.to_c()
        self.h = self.L.orc_create(ctypes.addressof(self._cs), self.S)
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.threads = int(threads) if threads else 1
        self.L.orc_set_threads(self.h, self.threads)

    def alloc_planes(self, n):
        return np.zeros((int(n), self.S), dtype=np.float64)

    def reset(self, vol0=None):
        self.L.orc_reset(self.h, _dp(vol0))

    def set_consts(self, consts):
        self.L.orc_set_consts(self.h, _dp(np.ascontiguousarray(consts, dtype=np.float64)))

    def step(self, days, slots, node_flow=None, col_flow=None, objective=None):
        return self.L.orc_step(self.h, float(days), _dp(slots), _dp(node_flow), _dp(col_flow), _dp(objective))

    def status(self):
        a, b = ctypes.c_int32(), ctypes.c_int32()
        self.L.orc_status(self.h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def get_basis(self):
        s = self.structure
        row = np.zeros((self.S, s.n_rows), np.int32)
        col = np.zeros((self.S, s.n_cols), np.int32)
        ip = ctypes.POINTER(ctypes.c_int32)
        self.L.orc_get_basis(self.h, row.ctypes.data_as(ip), col.ctypes.data_as(ip))
        return row, col

    def counters(self):
        v = [ctypes.c_int64() for _ in range(3)]
        self.L.orc_counters(self.h, *[ctypes.byref(x) for x in v])
        return {"pivots": v[0].value, "refactors": v[1].value, "steps": v[2].value}

    # -- program mode (device-resident run semantics) ------------------------------------------
    def load_program(self, prog):
        self._cp = prog.to_c()
        self._prog = prog
        rc = self.L.orc_load_program(self.h, ctypes.addressof(self._cp))
        if rc != 0:
            raise RuntimeError(f"orc_load_program failed ({rc})")

    def program_reset(self):
        self.L.orc_program_reset(self.h)

    def update_table(self, table, data):
        data = np.ascontiguousarray(data, dtype=np.float64)
        if self.L.orc_update_table(self.h, int(table), _dp(data)) != 0:
            raise RuntimeError("orc_update_table: bad table index or no program")

    def run(self, t0, n):
        rc = self.L.orc_run(self.h, int(t0), int(n))
        if rc < 0:
            raise RuntimeError("orc_run: bad range or no program")
        return rc

    def sync(self):
        pass

    def program_status(self):
        v = [ctypes.c_int32() for _ in range(4)]
        self.L.orc_program_status(self.h, *[ctypes.byref(x) for x in v])
        return tuple(x.value for x in v)

    def fetch_recorder(self, r, shape):
        out = np.zeros(shape, dtype=np.float64)
        if self.L.orc_fetch_recorder(self.h, int(r), _dp(out)) != 0:
            raise RuntimeError("orc_fetch_recorder failed")
        return out

    def fetch_recorder_agg(self, r):
        """(sum, min, max) over the timesteps of array recorder ``r`` (the CUDA engine accumulates them on the device)."""
        a = self.fetch_recorder(r, (self._prog.n_steps, self.S))
        # cumsum adds the rows one after the other, like the device (np.sum may reorder / pair the additions)
        return np.cumsum(a, axis=0)[-1], a.min(axis=0), a.max(axis=0)

    def fetch_state(self):
        s = self.structure
        x = np.zeros((s.n_cols, self.S))
        vol = np.zeros((max(s.n_storage, 1), self.S))
        pc = np.zeros((max(s.n_storage, 1), self.S))
        self.L.orc_fetch_state(self.h, _dp(x), _dp(vol), _dp(pc))
        return x, vol, pc

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None
