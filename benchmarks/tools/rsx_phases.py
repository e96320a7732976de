"""Phase cycles of rsx_solve_kernel (debug build: nvcc -DRSX_PHASE_TIMING -o pywr_b200/libb200lp_phase.so, see
benchmarks/tools/exp.sh) on the config-4 fixture: RSX_FIXTURE=network600|netspec2000 python benchmarks/tools/rsx_phases.py [S] [steps] [mode]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["B200LP_LIB"] = os.path.join(ROOT, "pywr_b200", "libb200lp_phase.so")
os.environ["B200LP_LARGE_ENGINE"] = "rsx"
if len(sys.argv) > 3:
    os.environ["B200LP_RSX_MODE"] = sys.argv[3]
import numpy as np
from pywr_b200.engine import CudaEngine
from pywr_b200.structure import LPStructure
FIX = os.path.join(ROOT, "benchmarks", "fixtures", os.environ.get("RSX_FIXTURE", "network600") + ".edge")
s = LPStructure.load(FIX + ".structure.npz"); tr = np.load(FIX + ".trace.npz")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng = CudaEngine(s, S, device=0)
slots = eng.alloc_planes(s.n_slots)
eng.reset(None)
names = ["-", "load_bounds", "load_state", "place_nonbasics", "g = N x_N", "beta = -W g", "leaving row / exit", "write_solution+residual", "store_state",
         "pivot: column w + check", "pivot: d / beta", "pivot: update W", "pivot: row rho", "pivot: alpha", "pivot: ratio tests"]
prev = np.zeros(18, np.uint64)
for t in range(T):
    base = tr["slots"][t]
    slots[:] = base[:, np.arange(S) % base.shape[1]]
    eng.step(float(tr["days"][t]), slots, None, None, None)
    out = (ctypes.c_ulonglong * 18)()
    eng.L.b200lp_debug_rsx_phases.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]
    assert eng.L.b200lp_debug_rsx_phases(eng.h, out) == 0
    cur = np.array(list(out), np.uint64)
    d = (cur - prev).astype(float); prev = cur
    tot = d[2:17].sum()
    print(f"t={t} pivots={int(d[0])} cycles/scenario={tot / S:.0f}: " + ", ".join(f"{names[k]} {100 * d[2 + k] / tot:.1f}%" for k in range(1, 15)))
    if d[0] > 0:
        print("   cycles per pivot: " + ", ".join(f"{names[k]} {d[2 + k] / d[0]:.0f}" for k in (6, 12, 13, 14, 9, 10, 11)))
