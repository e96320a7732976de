#!/usr/bin/env python
"""bench.py — scenario-timesteps/s of the B200 scenario-batched LP engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU path (oracle port, all host cores)

Workload (config.workload): BASELINE.json configs[1] — two_reservoir, 1,024 synthetic stochastic
inflow scenarios x 50 years daily (T = 18,262), per GPU (weak scaling: every rank runs its own
contiguous block of 1,024 scenarios; scenarios never exchange data, so no collective on the data
path; NCCL is used once after the timed region to gather the per-scenario recorder values).

A "step" is one whole run of the workload: all T timesteps for all scenarios of the rank through the
device-resident run (parameters -> bounds/costs -> warm-started simplex -> flow commit -> storage mass
balance -> recorders, one kernel launch).
  value : inputs resident in HBM when the timed region starts (program loaded, inflow table uploaded);
          timed with CUDA events on the engine's stream, max over ranks.
  e2e   : the same run through the C-ABI with HOST buffers (b200lp_run_pipelined): every step uploads the
          inflow table from pinned host memory, resets, runs and reads every recorder back to the host, the
          copies overlapped with the solve in 16 time slices; wall clock around the step, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "scenario-timesteps/sec"
UNIT = "scenario-timesteps/s"
WORKLOAD = "two_reservoir"
S_PER_GPU = 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--scenarios", type=int, default=S_PER_GPU, help="scenarios per GPU")
    ap.add_argument("--years", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=2.0):
        """nvidia-smi needs a few hundred ms to start: do not begin a short timed region before it samples."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, window=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        lines = [ln for ts, ln in self.lines if window is None or window[0] - 0.05 <= ts <= window[1] + 0.15]
        if not lines:  # timed region shorter than the sampling period: the samples around it (warm-up, also under load)
            lines = [ln for _, ln in self.lines]
        self.lines = lines
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu(s, prog, info, steps, warmup, threads):
    """Times the CPU path: the oracle (C restatement of the reference's per-scenario loop plus
    Storage.after and the recorders), all host threads, whole workload per step."""
    from oracle.oracle_engine import OracleEngine

    eng = OracleEngine(s, info["n_scenarios"], threads=threads)
    eng.load_program(prog)
    times = []
    for i in range(warmup + steps):
        eng.program_reset()
        t0 = time.perf_counter()
        eng.run(0, prog.n_steps)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    nfail = eng.program_status()[0]
    eng.close()
    units = info["n_scenarios"] * prog.n_steps
    return units * len(times) / sum(times), float(np.mean(times)), nfail


E2E_CHUNKS = int(os.environ.get("B200_BENCH_CHUNKS", "16"))   # time slices of the pipelined end-to-end run


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from benchmarks.workloads import algorithmic_bytes_per_scenario_step, load_workload

    S = args.scenarios
    config = {"workload": f"{args.workload}: {S} synthetic stochastic inflow scenarios per GPU x "
                          f"{args.years or 50 if args.workload == 'two_reservoir' else args.years or 30} years daily",
              "scenarios_per_gpu": S, "formulation": "route", "kernel": "register-resident dictionary simplex, "
              "device-resident run", "l2": "inputs+outputs per step (inflow table + recorder arrays) exceed the 126 MB L2",
              "parallelism": f"scenario-sharded x{world}, no data-path collective"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        s, prog, info = load_workload(args.workload, S, years=args.years)
        config.update(timesteps=info["timesteps"], lp_rows=info["rows"], lp_cols=info["cols"])
        cores = cpu_cores()
        value, mean_t, nfail = run_cpu(s, prog, info, args.steps, max(args.warmup, 1), cores)
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"whole workload per step ({S} scenarios x {info['timesteps']} timesteps); "
                                       "GLPK is not installable in this image, so the CPU arm is the oracle port "
                                       "(oracle/lp_oracle.c, program mode, pthreads over scenario blocks)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "failed_scenarios": nfail,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    from pywr_b200.engine import CudaEngine

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep rank 0's stdout to the one JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

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

    s, prog, info = load_workload(args.workload, S, years=args.years, first_scenario=rank * S)
    T = info["timesteps"]
    config.update(timesteps=T, lp_rows=info["rows"], lp_cols=info["cols"], recorders=info["n_recorders"])
    eng = CudaEngine(s, S, device=local_rank)
    eng.load_program(prog)
    inflow_table = int(np.nonzero(prog.table_axis >= 0)[0][0])
    units = S * T

    # ---- value: inputs resident in HBM, CUDA events on the engine's stream -------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        eng.program_reset()
        eng.run(0, T)
    eng.sync()
    st0 = eng.stats()
    sampler.wait_first()
    barrier()
    kernel_ms = []
    eng.sync()
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        eng.program_reset()           # restores initial volumes / bases / zeroed recorders (not timed below)
        eng.mark(0)
        eng.run(0, T)
        eng.mark(1)
        kernel_ms.append(eng.elapsed_ms())
    eng.sync()
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop((t_wall0, t_wall0 + wall))
    st1 = eng.stats()
    nfail = eng.program_status()[0]
    dev_time = max_over_ranks(sum(kernel_ms) / 1e3)
    value = world * units * args.steps / dev_time
    launches = int(st1["kernel_launches"] - st0["kernel_launches"])

    # ---- e2e: host buffers in, host buffers out, every step -----------------------------------
    e2e = None
    if not args.no_e2e:
        r0, c0 = int(prog.table_rows[inflow_table]), int(prog.table_cols[inflow_table])
        host_in = eng.alloc_pinned((r0, c0))
        off = int(prog.table_offset[inflow_table])
        host_in[...] = prog.table_data[off: off + r0 * c0].reshape(r0, c0)
        host_out = [eng.alloc_pinned(prog.rec_shape(r, S)) for r in range(prog.n_recorders)]
        h2d = host_in.nbytes
        d2h = sum(a.nbytes for a in host_out)

        def e2e_step():
            # the call a user makes: host table in, host recorder arrays out, time slices pipelined
            eng.run_pipelined({inflow_table: host_in}, host_out, chunks=E2E_CHUNKS)

        for _ in range(2):
            e2e_step()
        eng.sync()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        eng.sync()
        barrier()
        e2e_time = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * units * args.steps / e2e_time, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_time / args.steps * 1e3,
               "api": f"b200lp_run_pipelined: reset + upload + run + read-out of every recorder, {E2E_CHUNKS} time slices on three "
                      "streams (pinned host buffers)"}

    # ---- multi-GPU read-out: NCCL gather of the per-scenario recorder values -----------------
    gather_ms = None
    if dist is not None:
        import torch

        class _Dev:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

        t0 = time.perf_counter()
        for r in range(prog.n_recorders):
            if prog.rec_shape(r, S) == (S,):
                ptr, n = eng.recorder_device_ptr(r)
                local = torch.as_tensor(_Dev(ptr, n), device=f"cuda:{local_rank}")
                out = torch.empty(world * n, dtype=torch.float64, device=f"cuda:{local_rank}")
                dist.all_gather_into_tensor(out, local)
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - t0) * 1e3

    # ---- roofline of the dominant (only) kernel ---------------------------------------------------
    peak, peak_src = measured_hbm_peak()
    b_alg = algorithmic_bytes_per_scenario_step(info)
    mean_kernel_s = float(np.mean(kernel_ms)) / 1e3
    achieved = b_alg * units / mean_kernel_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("lp_run_kernel_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "lp_run_kernel",
                "algorithmic_bytes_per_scenario_timestep": b_alg, "units_per_launch": units,
                "note": "the kernel is issue/latency bound (no dense contraction, per-scenario state in registers); "
                        "HBM fraction is reported as the contract asks"}

    # ---- CPU baseline on rank 0 at N=1 ------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = cpu_cores()
        v, mean_t, _ = run_cpu(s, prog, info, 1, 1, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"whole workload once ({S} x {T}); oracle/lp_oracle.c program mode on {cores} threads "
                         "(GLPK is not installable in this image)"}

    # ---- the same engine driven by the reference framework itself (only where Pywr is importable) ----
    pywr_extra = None
    if world == 1 and not args.no_e2e:
        try:
            ref_dir = os.path.join(ROOT, "baseline", "_ref")
            if os.path.isdir(os.path.join(ref_dir, "pywr")) and ref_dir not in sys.path:
                sys.path.insert(0, ref_dir)
            import pywr_b200

            pywr_b200.register()  # the package was imported before Pywr was importable
            from benchmarks.workloads import build_model
            from pywr_b200.device_run import run_on_device

            yrs = 2
            m1 = build_model(args.workload, S, years=yrs, solver="b200")
            res1 = m1.run()
            steps1 = (res1.timestep.index + 1) * S
            m2 = build_model(args.workload, S, years=yrs, solver="null")
            res2 = run_on_device(m2)
            pywr_extra = {
                "model_run_solver_b200": {"value": steps1 / res1.time_taken, "unit": UNIT, "years": yrs,
                                          "note": "Model.run(): Pywr's own before()/after() on the host every timestep, "
                                                  "one b200lp_step per timestep"},
                "run_on_device": {"value": res2.speed, "unit": UNIT, "years": yrs, "compile_s": res2.time_compile,
                                  "readout_s": res2.time_readout,
                                  "note": "pywr_b200.device_run.run_on_device(model): whole run in one kernel launch, "
                                          "recorders written back into Pywr's objects"},
            }
        except Exception as exc:  # Pywr not available on this box
            pywr_extra = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_time / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "failed_scenarios": nfail, "wall_ms_per_step_including_reset": wall / args.steps * 1e3,
            "pivots_per_scenario_timestep": (st1["pivots"] - st0["pivots"]) / max(1, units * args.steps),
            "kernel_variant": st1["kernel_variant"], "gather_ms": gather_ms, "through_pywr": pywr_extra,
        }
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
