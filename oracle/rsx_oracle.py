"""Host build of the revised-simplex core (pywr_b200/csrc/rsx_core.cuh) behind the engine interface.
TEST INFRASTRUCTURE ONLY (same rule as oracle/oracle_engine.py).

The reduced, scaled LP of every scenario comes from the numpy restatement of the PDHG front end
(``oracle.pdhg_oracle.PdhgPlan`` / ``assemble``: structural presolve of the singleton rows of
cython_glpk.pyx:1300-1312, Ruiz scaling); the solve is the SAME source file the CUDA library compiles for the
device, instantiated with a one-thread context (oracle/rsx_host.cpp).  Scenario 0 of the first timestep starts
from the slack basis and its basis / inverse seed the other scenarios, exactly like the device driver
(pywr_b200/csrc/pdhg_host.inc::rsx_solve).
"""
import ctypes
import os
import subprocess

import numpy as np

from .pdhg_oracle import ST_OPTIMAL, PdhgOracleEngine, assemble, snap_unscale

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "librsxhost.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", _HERE, "build/librsxhost.so"])
        L = ctypes.CDLL(_LIB_PATH)
        dp, ip, bp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8)
        L.rsx_host_run.restype = ctypes.c_int
        L.rsx_host_run.argtypes = [ctypes.c_int, ctypes.c_int, ip, ip, dp, ip, ip, dp, ctypes.c_int, dp, ip, bp, dp, ip,
                                   dp, dp, dp, dp, dp, dp, ctypes.c_int, ip, ip, ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


class RsxOracleEngine(PdhgOracleEngine):
    def __init__(self, structure, n_scenarios, mode=0, refresh_pivots=1000, **kw):
        self.refresh_pivots = refresh_pivots   # rebuild the reduced costs after this many pivots of a scenario (device: 1000)
        self.mode = mode   # placement of the working vectors (rsx_core.cuh::work_carve); no effect on the arithmetic
        super().__init__(structure, n_scenarios, **kw)
        self._bind()

    def _bind(self):
        p = self.plan
        A = p.A.tocsr(); A.sort_indices()
        AT = p.A.tocsc(); AT.sort_indices()
        self.csr = (A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(float))
        self.csc = (AT.indptr.astype(np.int32), AT.indices.astype(np.int32), AT.data.astype(float))
        m, n = p.m_r, p.n
        self.ldw = (m + 1) & ~1
        S = self.S
        self.W = np.zeros((S, max(m, 1) * max(self.ldw, 1)))
        self.head = np.zeros((S, max(m, 1)), np.int32)
        self.stat = np.zeros((S, n + m), np.uint8)
        self.d = np.zeros((S, n + m))
        self.info = np.zeros((S, 4), np.int32)
        self.xs = np.zeros((S, n))
        self.pivots = np.zeros(1, np.int32)
        self.refactors = np.zeros(1, np.int32)
        self.seeded = False
        s = self.structure
        self.cost_dynamic = int(any(r >= 0 for r in s.cost_ref[: s.cost_indptr[s.n_cols]]))

    def reset(self, vol0=None):
        super().reset(vol0)
        if hasattr(self, "info"):
            self.info[:] = 0
            self.seeded = False

    def set_consts(self, consts):
        super().set_consts(consts)
        self._bind()

    def _solve(self, k, lo, hi, c, lc, uc, cold):
        p = self.plan
        L = lib()
        d, i, b = ctypes.c_double, ctypes.c_int32, ctypes.c_uint8
        arrs = [np.ascontiguousarray(a, float) for a in (c, lc, uc, lo, hi)]
        return L.rsx_host_run(p.m_r, p.n, _p(self.csr[0], i), _p(self.csr[1], i), _p(self.csr[2], d), _p(self.csc[0], i),
                              _p(self.csc[1], i), _p(self.csc[2], d), self.cost_dynamic, _p(self.W[k], d), _p(self.head[k], i),
                              _p(self.stat[k], b), _p(self.d[k], d), _p(self.info[k], i), *[_p(a, d) for a in arrs],
                              _p(self.xs[k], d), int(cold), _p(self.pivots, i), _p(self.refactors, i), int(self.mode), int(self.refresh_pivots))

    def step(self, days, slots, node_flow=None, col_flow=None, objective=None):
        p = self.plan
        nbad = 0
        data = [assemble(p, slots[:, k], days, self.vol[:, k]) for k in range(self.S)]
        if not self.seeded:
            if data[0][5] == ST_OPTIMAL:
                lo, hi, c, lc, uc, _ = data[0]
                self._solve(0, lo, hi, c, lc, uc, cold=True)
                for k in range(1, self.S):
                    self.W[k] = self.W[0]; self.head[k] = self.head[0]; self.stat[k] = self.stat[0]
                    self.d[k] = self.d[0]; self.info[k] = self.info[0]
            self.seeded = True
        for k in range(self.S):
            lo, hi, c, lc, uc, status = data[k]
            if status == ST_OPTIMAL:
                status = self._solve(k, lo, hi, c, lc, uc, cold=False)
            self._status[k] = status
            if status != ST_OPTIMAL:
                nbad += 1
                continue
            x = snap_unscale(p, self.xs[k], lc, uc)
            if col_flow is not None:
                col_flow[:, k] = x
            if node_flow is not None:
                node_flow[:, k] = self.flowmat @ x
            if objective is not None:
                objective[0, k] = (c / p.dc) @ x
        return nbad

    def stats(self):
        return {"pivots": int(self.pivots[0]), "refactors": int(self.refactors[0]), "n_rows": self.structure.n_rows,
                "n_cols": self.structure.n_cols, "reduced_rows": self.plan.m_r}
