"""the generated ~2,000-node network (edge formulation, LP 3,994 x 2,672)
on the batched restarted-PDHG path, S scenarios x T daily steps replayed from the committed fixture
(tests/golden/make_network_fixture.py).  Scenario k scales every recorded input plane by one
log-normal factor (scenario 0: factor 1, so its objective must equal HiGHS's recorded optimum).

Prints one JSON line: scenario-timesteps/s, PDHG iterations, and the achieved HBM bandwidth of the
iteration kernels against the algorithmic bytes per scenario-iteration  8 (7 n + 5 m_r + 2 nnz_r).

usage: python benchmarks/pdhg_bench.py [--scenarios 8192] [--steps 12] [--tol 2.5e-8]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pywr_b200.engine import CudaEngine  # noqa: E402
from pywr_b200.structure import LPStructure  # noqa: E402

FIX = os.path.join(ROOT, "benchmarks", "fixtures", "network600.edge")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--tol", type=float, default=None)
    ap.add_argument("--seed", type=int, default=20240903)
    ap.add_argument("--tail", type=int, default=None, help="pdhg_tail option: scenarios handed to the CTA-per-scenario kernel")
    ap.add_argument("--engine", choices=["auto", "pdhg", "rsx"], default="pdhg",
                    help="method of the large-LP path: pdhg (iterations), rsx (revised simplex), auto (what b200lp_create picks)")
    a = ap.parse_args()
    if a.engine != "auto":
        os.environ["B200LP_LARGE_ENGINE"] = a.engine
    s = LPStructure.load(FIX + ".structure.npz")
    tr = np.load(FIX + ".trace.npz")
    S, T = a.scenarios, min(a.steps, len(tr["days"]))
    rng = np.random.default_rng(a.seed)
    factor = np.exp(rng.normal(0.0, 0.25, size=S))
    factor[0] = 1.0
    vol_slots = sorted({int(r) for r in s.st_vol_ref if r >= 0})
    eng = CudaEngine(s, S, device=0)
    st0 = eng.stats()
    assert st0["kernel_variant"] in (200, 300), "expected the large-LP path"
    rsx = st0["kernel_variant"] == 300
    if a.tol is not None:
        eng.set_option("pdhg_tol", a.tol)
    if a.tail is not None:
        eng.set_option("pdhg_tail", a.tail)
    slots = eng.alloc_planes(s.n_slots)
    nf = eng.alloc_planes(s.n_nodes)
    ob = eng.alloc_planes(1)
    eng.reset(None)
    rows = []
    t_all = time.perf_counter()
    for t in range(T):
        base = tr["slots"][t][:, 0]
        slots[:] = np.where(np.isfinite(base), base, 0.0)[:, None] * factor[None, :]
        inf = ~np.isfinite(base)
        slots[inf] = base[inf][:, None]
        slots[vol_slots] = base[vol_slots][:, None]   # volumes are state, not scaled
        before = eng.stats()
        w0 = time.perf_counter()
        nbad = eng.step(float(tr["days"][t]), slots, nf, None, ob)
        wall = time.perf_counter() - w0
        after = eng.stats()
        ref = float(tr["objective_highs"][t, 0])
        rows.append({"t": t, "wall_s": wall, "device_ms": after["device_ms"] - before["device_ms"],
                     "iters": after["pdhg_iterations"] - before["pdhg_iterations"], "pivots": after["pivots"] - before["pivots"],
                     "failed": int(nbad),
                     "obj_rel_err_vs_highs": abs(float(ob[0, 0]) - ref) / max(1.0, abs(ref))})
        if nbad:
            bad, code = eng.status()
            rows[-1].update(first_bad=bad, code=code, factor=float(factor[bad]))
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    total_wall = time.perf_counter() - t_all
    st = eng.stats()
    n, m_r, nnz_r = s.n_cols, st["reduced_rows"], st["reduced_nonzero"]
    bytes_iter = 8 * (7 * n + 5 * m_r + 2 * nnz_r)
    dev_s = sum(r["device_ms"] for r in rows) / 1e3
    # every scenario runs until ALL of its batch has converged only in launch count, not in traffic:
    # converged scenario pairs return before touching memory, so traffic follows scenario-iterations
    it_total = st["pdhg_iterations"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6541.8)) if isinstance(peaks, dict) else 6541.8
    out = {"metric": "scenario-timesteps/sec (PDHG path, config 4)", "value": S * T / dev_s, "unit": "scenario-timesteps/s",
           "e2e_value": S * T / total_wall, "scenarios": S, "steps": T, "rows": s.n_rows, "cols": n, "nnz": s.nnz,
           "reduced_rows": m_r, "reduced_nnz": nnz_r, "tol": a.tol if a.tol is not None else 1e-8,
           "iterations_per_scenario_step": it_total / (S * T), "first_step_iterations_per_scenario": rows[0]["iters"] / S,
           "warm_iterations_per_scenario_step": (it_total - rows[0]["iters"]) / max(1, S * (T - 1)),
           "device_s": dev_s, "wall_s": total_wall, "failed": int(sum(r["failed"] for r in rows)),
           "max_obj_rel_err_vs_highs_scenario0": max(r["obj_rel_err_vs_highs"] for r in rows),
           # HBM roofline of the streaming kernels: only meaningful when every iteration ran on them (--tail 0);
           # the CTA-per-scenario kernel keeps the vectors in shared memory and moves no HBM bytes per iteration
           "roofline": ({"bound": "hbm", "achieved": bytes_iter * it_total / dev_s / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_iter * it_total / dev_s / 1e9 / peak, "bytes_per_scenario_iteration": bytes_iter}
                        if a.tail == 0 else None),
           "engine": "streaming kernels only" if a.tail == 0 else "streaming first batch, then CTA-per-scenario shared-memory kernel",
           "kernel_launches": st["kernel_launches"], "h2d_bytes": st["h2d_bytes"], "d2h_bytes": st["d2h_bytes"]}
    if rsx:
        # revised simplex: HBM traffic is dominated by the one full read of the scenario's basis inverse per
        # scenario-timestep (beta = -W g); pivots add rows / columns / the sparse rank-1 update
        ldw = (m_r + 1) & ~1
        wbytes = 8 * m_r * ldw
        for k in ("iterations_per_scenario_step", "first_step_iterations_per_scenario", "warm_iterations_per_scenario_step"):
            out.pop(k)
        warm = rows[1:] if len(rows) > 1 else rows
        warm_s = sum(r["device_ms"] for r in warm) / 1e3
        out.update({"metric": "scenario-timesteps/sec (revised simplex, config 4)", "engine": "revised simplex, explicit basis inverse in HBM, one CTA per scenario",
                    "pivots": st["pivots"], "refactors": st["refactors"], "pivots_per_scenario_step": st["pivots"] / (S * T),
                    "first_step_device_s": rows[0]["device_ms"] / 1e3,
                    "warm_value": S * len(warm) / warm_s if warm_s > 0 else None,
                    "roofline": {"bound": "hbm", "achieved": wbytes * S * len(warm) / warm_s / 1e9 if warm_s > 0 else None, "peak": peak, "unit": "GB/s",
                                 "frac": (wbytes * S * len(warm) / warm_s / 1e9 / peak) if warm_s > 0 else None,
                                 "bytes_per_scenario_timestep": wbytes, "note": "warm timesteps; algorithmic bytes = one read of W per scenario-timestep"}})
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
