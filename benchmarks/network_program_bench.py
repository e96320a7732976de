"""SURVEY.md 8(d) config 4 as a DEVICE-RESIDENT run: the generated 2,073-node network (edge formulation, LP 3,994 x 2,672,
600 catchment inflows = monthly profile x per-scenario wetness) for S scenarios x T daily timesteps — parameters, bounds,
warm-started revised simplex, flow commit, storage mass balance and recorders all on the GPU (b200lp_load_program +
b200lp_run on the large-LP path: plane kernels around the batched solve), results fetched at the end.

Scenarios 0 and 1 are the two worlds of the committed fixture (tests/golden/make_network_program.py), whose recorder arrays
come from a host run of the reference framework: they are checked here at full size.  Prints one JSON line.

usage: python benchmarks/network_program_bench.py [--scenarios 8192] [--steps 365] [--engine auto|rsx|pdhg]
N GPUs of one box (weak scaling: --scenarios per GPU, scenario shards with their own wetness draws, no data-path collective;
NCCL only for the barrier and the max-over-ranks time):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P benchmarks/network_program_bench.py
"""
import argparse
import copy
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pywr_b200.engine import CudaEngine  # noqa: E402
from pywr_b200.program import Program  # noqa: E402
from pywr_b200.structure import LPStructure  # noqa: E402

FIXTURES = {
    # round-1 network: 2,073 nodes, one licence node, monthly-profile inflows x a per-scenario constant
    "network600": os.path.join(ROOT, "benchmarks", "fixtures", "network600"),
    # SURVEY 8(d) config 4 as specified: 40 aggregated nodes (20 with factors), 60 piecewise links, cycles, per-scenario DAILY
    # AR(1) inflows (benchmarks/workloads.py::spec_network_document, tests/golden/make_netspec.py full)
    "netspec2000": os.path.join(ROOT, "benchmarks", "fixtures", "netspec2000"),
}
FIX = FIXTURES["network600"]


def widen(s, base, S, T, seed=20240903, rank=0):
    """S scenarios: the per-scenario table(s) (the wetness factor) get S columns; scenario k < 2 keeps the fixture's value.
    Time-indexed per-scenario tables (netspec2000: the regional daily AR(1) series) are regenerated for the global scenario
    ids of this rank with the generator that made the fixture (benchmarks.workloads.regional_wetness)."""
    from benchmarks.workloads import regional_wetness

    rng = np.random.default_rng([seed, rank])
    q = copy.copy(base)
    q.n_steps = T
    q.ts_days = base.ts_days[:T].copy()
    rows, cols, offs, chunks, pos = [], [], [], [], 0
    for t in range(base.n_tables):
        r0, c0 = int(base.table_rows[t]), int(base.table_cols[t])
        data = base.table_data[base.table_offset[t]: base.table_offset[t] + r0 * c0].reshape(r0, c0)
        if base.table_axis[t] >= 0 and r0 > 1:
            region = next(r for r in range(16) if np.array_equal(regional_wetness(r0, 1, r)[:, 0], data[:, 0]))
            data = regional_wetness(r0, S, region, first_scenario=rank * S)[:T]
        elif base.table_axis[t] >= 0:
            wide = np.repeat(data[:, :1], S, axis=1) * rng.uniform(0.5, 1.25, size=(1, S))
            wide[:, : min(c0, S)] = data[:, : min(c0, S)]
            data = wide
        rows.append(data.shape[0]); cols.append(data.shape[1]); offs.append(pos)
        chunks.append(np.ascontiguousarray(data).reshape(-1)); pos += data.size
    q.table_rows, q.table_cols = np.array(rows, np.int32), np.array(cols, np.int32)
    q.table_offset = np.array(offs, np.int64)
    q.table_data = np.concatenate(chunks)
    q.scen_index = np.tile(np.arange(S, dtype=np.int32), (max(base.n_axes, 1), 1)).reshape(-1)
    nst = max(s.n_storage, 1)
    q.initial_volume = np.repeat(base.initial_volume.reshape(nst, -1)[:, :1], S, axis=1).reshape(-1)
    q.initial_pc = np.repeat(base.initial_pc.reshape(nst, -1)[:, :1], S, axis=1).reshape(-1)
    return q


def check_against_fixture(outs, prog, T, full_T):
    """scenarios 0 and 1 against the reference framework's host run: array recorders over the first T timesteps, the
    accumulating ones only when the whole fixture run was executed"""
    z = np.load(FIX + ".expected.npz")
    names = [str(n) for n in z["names"]]
    worst, checked = 0.0, 0
    for i in range(prog.n_recorders):
        exp, out = z[f"rec{i}"], outs[i]
        if out.ndim == 2:
            a, b = out[:T, :2], exp[:T, :2]
        elif T == full_T:
            a, b = out[:2], exp[:2]
            if names[i].startswith(("mf_", "df_")):   # the engine accumulates; the reference divides by T at read-out
                a = a / T
        else:
            continue
        d = float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))
        worst, checked = max(worst, d), checked + 1
    return worst, checked


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=365)
    ap.add_argument("--engine", choices=["auto", "pdhg", "rsx"], default="auto")
    ap.add_argument("--fixture", choices=sorted(FIXTURES), default="network600")
    ap.add_argument("--phases", action="store_true", help="per-phase cycles of rsx_solve_kernel (needs B200LP_LIB = a -DRSX_PHASE_TIMING build)")
    a = ap.parse_args()
    global FIX
    FIX = FIXTURES[a.fixture]
    if a.engine != "auto":
        os.environ["B200LP_LARGE_ENGINE"] = a.engine
    world, rank, local_rank = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # one JSON line on stdout
        torch.cuda.set_device(local_rank)
        import ctypes

        sys.stdout.flush()
        saved_fd = os.dup(1)          # NCCL's version banner (C stdio, rank 0) goes to stderr: stdout carries the JSON line only
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
            ctypes.CDLL(None).fflush(None)
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    s = LPStructure.load(FIX + ".pstructure.npz")
    base = Program.load(FIX + ".program.npz")
    S, T = a.scenarios, min(a.steps, base.n_steps)
    q = widen(s, base, S, T, rank=rank)
    eng = CudaEngine(s, S, device=local_rank)
    st0 = eng.stats()
    w0 = time.perf_counter()
    eng.load_program(q)
    load_s = time.perf_counter() - w0
    eng.sync()
    barrier()
    w0 = time.perf_counter()
    eng.mark(0)
    eng.run(0, T)
    eng.mark(1)
    eng.sync()
    barrier()
    wall = max_over_ranks(time.perf_counter() - w0)
    dev_s = max_over_ranks(eng.elapsed_ms() / 1e3)
    status = eng.program_status()
    w0 = time.perf_counter()
    outs = [eng.fetch_recorder(r, q.rec_shape(r, S)) for r in range(q.n_recorders)]
    read_s = time.perf_counter() - w0
    worst, checked = check_against_fixture(outs, q, T, base.n_steps) if rank == 0 else (0.0, 0)
    st = eng.stats()
    failed_all = int(max_over_ranks(float(status[0])))
    S_total = S * world
    m_r = st["reduced_rows"]
    wbytes = 8 * m_r * ((m_r + 1) & ~1)
    peak = 6541.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
    except Exception:
        pass
    out = {"metric": "scenario-timesteps/sec (config 4, device-resident run)", "fixture": a.fixture, "value": S_total * T / dev_s, "unit": "scenario-timesteps/s",
           "n_gpus": world, "scaling": "weak", "scenarios": S_total, "scenarios_per_gpu": S, "timesteps": T, "device_s": dev_s, "wall_s": wall, "load_program_s": load_s, "readout_s": read_s,
           "kernel_variant": st0["kernel_variant"], "rows": s.n_rows, "cols": s.n_cols, "reduced_rows": m_r, "nodes": s.n_nodes,
           "ops": int(q.n_ops), "tables": int(q.n_tables), "recorders": int(q.n_recorders), "failed_scenarios_max_over_ranks": failed_all,
           "pivots": st["pivots"], "refactors": st["refactors"], "pivots_per_scenario_timestep": st["pivots"] / (S * T),
           "pdhg_iterations": st["pdhg_iterations"], "kernel_launches": st["kernel_launches"],
           "max_rel_diff_vs_reference_host_run_scenarios_0_1": worst, "recorders_checked": checked,
           "roofline": {"bound": "hbm", "achieved": wbytes * S * T / dev_s / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": wbytes * S * T / dev_s / 1e9 / peak, "bytes_per_scenario_timestep": wbytes,
                        "note": "per GPU; algorithmic bytes = one read of the scenario's basis inverse per scenario-timestep"} if st0["kernel_variant"] == 300 else None}
    # SURVEY 8(d)'s algorithmic bytes of one scenario-timestep of the ORIGINAL LP: bounds of rows and columns, statuses, storage state,
    # node flows, recorder samples = 8 (n + 2 m) + 2 (m + n) + 16 n_storage + 8 N + 8 n_rec
    alg = 8 * (s.n_cols + 2 * s.n_rows) + 2 * (s.n_rows + s.n_cols) + 16 * s.n_storage + 8 * s.n_nodes + 8 * int(q.n_recorders)
    out["roofline_algorithmic"] = {"bound": "hbm", "bytes_per_scenario_timestep": alg, "achieved": alg * S * T / dev_s / 1e9, "peak": peak,
                                   "unit": "GB/s", "frac": alg * S * T / dev_s / 1e9 / peak,
                                   "note": "per GPU, against SURVEY 8(d)'s per-unit bytes of the LP itself; the engine's own stream (the dense "
                                           "basis inverse, `roofline`) is an implementation cost on top of these"}
    if a.phases and rank == 0:
        import ctypes

        names = ["-", "load_bounds", "load_state", "place_nonbasics", "g = N x_N", "beta = -W g", "leaving row / exit", "write_solution+residual",
                 "store_state", "pivot: column w + check", "pivot: d / beta", "pivot: update W", "pivot: row rho", "pivot: alpha", "pivot: ratio tests"]
        buf = (ctypes.c_ulonglong * 18)()
        eng.L.b200lp_debug_rsx_phases.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]
        assert eng.L.b200lp_debug_rsx_phases(eng.h, buf) == 0
        d = [float(v) for v in buf]
        tot = sum(d[2:17])
        out["rsx_phases"] = {"cycles_per_scenario_timestep": tot / (S * T),
                             "share": {names[k]: d[2 + k] / tot for k in range(1, 15)},
                             "cycles_per_pivot": {names[k]: d[2 + k] / max(d[0], 1.0) for k in (6, 12, 13, 14, 9, 10, 11)}}
    if rank == 0:
        print(json.dumps(out))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
