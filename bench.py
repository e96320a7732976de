#!/usr/bin/env python
"""bench.py — scenario-timesteps/s of the B200 scenario-batched LP engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU path (oracle port, all host cores)

Workload (config.workload): BASELINE.json configs[2], the configuration north_star's target is quoted on —
examples/thames with its 5-member `demand` scenario sliced to [2, 3] and 4,096 synthetic stochastic inflow
scenarios x 30 years daily (T = 10,957).  Default `--scaling weak`: 4,096 scenarios PER GPU (every rank runs its
own contiguous block of global scenario ids; scenarios never exchange data, so no collective on the data path;
NCCL is used once after the timed region to gather the per-scenario recorder values).  `--scaling strong`: 4,096
scenarios in TOTAL, 4,096 / N per GPU (north_star: "a 4,096-scenario Thames-scale model on one 8 x B200 box");
under torchrun the default run also measures that split and reports it under "north_star_target".
`--workload two_reservoir --scenarios 1024` is BASELINE.json configs[1] (reported under "config2" at N = 1).
BASELINE.json configs[3] (the 2,000-node network, large-LP path) is measured by benchmarks/network_program_bench.py, which the
default N = 1 run executes once in its own process and reports under "config4".

A "step" is one whole run of the workload: all T timesteps for all scenarios of the rank through the
device-resident run (parameters -> bounds/costs -> warm-started simplex -> flow commit -> storage mass
balance -> recorders, one kernel launch).
  value : inputs resident in HBM when the timed region starts (program loaded, inflow table uploaded);
          timed with CUDA events on the engine's stream, max over ranks.
  e2e   : the same run through the C-ABI with HOST buffers (b200lp_run_pipelined): every step uploads the
          inflow table from pinned host memory, resets, runs and reads every recorder's per-scenario result back
          to the host — the [S] values of the accumulating recorders and, for the array recorders, their temporal
          aggregates (sum / min / max per scenario = Recorder.values(), accumulated on the device); the copies
          overlap the solve in 16 time slices; wall clock around the step, max over ranks.  "e2e_full_arrays" is
          the same with every [T, S] recorder array shipped to the host as well.
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
WORKLOAD = "thames"
S_DEFAULT = {"thames": 4096, "two_reservoir": 1024}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--scenarios", type=int, default=None, help="scenarios per GPU (weak) / in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
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


E2E_CHUNKS = int(os.environ.get("B200_BENCH_CHUNKS", "0"))   # time slices of the pipelined end-to-end run; 0: by input size


def e2e_chunks(table_bytes):
    """16 slices for the full-size tables (measured best at 150 - 360 MB), about one per 12 MB below that (a slice costs a
    launch and a pair of event waits: 512 thames scenarios, 45 MB: 4 slices 6.92 ms, 16 slices 7.26 ms)."""
    if E2E_CHUNKS > 0:
        return E2E_CHUNKS
    return int(max(2, min(16, round(table_bytes / 12e6))))


def years_of(workload, years):
    return years or (50 if workload == "two_reservoir" else 30)


def workload_text(workload, S, years, scaling, world):
    per = "per GPU" if scaling == "weak" else f"in total ({S // world} per GPU)"
    extra = ", `demand` scenario (5 members) sliced to [2, 3]" if workload == "thames" else ""
    return f"{workload}{extra}: {S} synthetic stochastic inflow scenarios {per} x {years_of(workload, years)} years daily, {scaling} scaling"


def traffic_of(workload, S, T):
    """DRAM bytes per launch of the run kernel from the committed ncu capture of exactly this launch (or None)."""
    try:
        entries = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("entries", {})
        e = entries.get(f"{workload}:{S}:{T}")
        return (int(e["dram_bytes_read"]) + int(e["dram_bytes_write"])) if e else None
    except Exception:
        return None


class Bench:
    """The B200 arm for one (workload, scenarios per rank): program resident on the device, timed runs."""

    def __init__(self, workload, S_rank, years, first_scenario, device):
        from benchmarks.workloads import load_workload
        from pywr_b200.engine import CudaEngine

        self.workload, self.S = workload, S_rank
        self.s, self.prog, self.info = load_workload(workload, S_rank, years=years, first_scenario=first_scenario)
        self.T = self.info["timesteps"]
        self.eng = CudaEngine(self.s, S_rank, device=device)
        self.eng.load_program(self.prog)
        self.inflow_table = int(np.nonzero(self.prog.table_axis >= 0)[0][0])
        self.units = S_rank * self.T

    def warm(self, n):
        for _ in range(n):
            self.eng.program_reset()
            self.eng.run(0, self.T)
        self.eng.sync()

    def timed_runs(self, steps):
        """kernel time of each of `steps` whole runs (CUDA events on the engine's stream)"""
        ms = []
        for _ in range(steps):
            self.eng.program_reset()      # restores initial volumes / bases / zeroed recorders (not timed)
            self.eng.mark(0)
            self.eng.run(0, self.T)
            self.eng.mark(1)
            ms.append(self.eng.elapsed_ms())
        self.eng.sync()
        return ms

    def e2e_setup(self, full_arrays):
        from pywr_b200.program import ARRAY_KINDS

        prog, eng, S = self.prog, self.eng, self.S
        r0, c0 = int(prog.table_rows[self.inflow_table]), int(prog.table_cols[self.inflow_table])
        if not hasattr(self, "host_in"):
            self.host_in = eng.alloc_pinned((r0, c0))
            off = int(prog.table_offset[self.inflow_table])
            self.host_in[...] = prog.table_data[off: off + r0 * c0].reshape(r0, c0)
        arrays = [r for r in range(prog.n_recorders) if int(prog.rec_kind[r]) in ARRAY_KINDS]
        outs = [eng.alloc_pinned(prog.rec_shape(r, S)) if (full_arrays or r not in arrays) else None
                for r in range(prog.n_recorders)]
        d2h = sum(a.nbytes for a in outs if a is not None) + (0 if full_arrays else 3 * 8 * S * len(arrays))

        chunks = self.chunks = e2e_chunks(self.host_in.nbytes)

        def step():
            # the call a user makes: host table in, per-scenario recorder results out, time slices pipelined
            eng.run_pipelined({self.inflow_table: self.host_in}, outs, chunks=chunks)
            if not full_arrays:
                for r in arrays:
                    eng.fetch_recorder_agg(r)     # (sum, min, max) over time per scenario: Recorder.values()
        return step, int(self.host_in.nbytes), int(d2h)

    def close(self):
        self.eng.close()


def single_process_run(workload, S_total, years, n_devices, steps):
    """S_total scenarios over n_devices GPUs driven by ONE process (pywr_b200.engine.ShardedEngine, what
    run_on_device(model, devices="all") uses): wall clock around reset-excluded runs, synchronised on both sides."""
    from benchmarks.workloads import load_workload
    from pywr_b200.engine import ShardedEngine

    s, prog, info = load_workload(workload, S_total, years=years)
    eng = ShardedEngine(s, S_total, devices=list(range(n_devices)))
    eng.load_program(prog)
    T = info["timesteps"]
    for _ in range(3):
        eng.program_reset(); eng.run(0, T)
    eng.sync()
    times = []
    for _ in range(steps):
        eng.program_reset()
        eng.sync()
        t0 = time.perf_counter()
        eng.run(0, T)
        eng.sync()
        times.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    outs = [eng.fetch_recorder_agg(r) if prog.rec_shape(r, S_total) != (S_total,) else eng.fetch_recorder(r, (S_total,))
            for r in range(prog.n_recorders)]
    readout = time.perf_counter() - t0
    nfail = eng.program_status()[0]
    eng.close()
    del outs
    return {"workload": workload_text(workload, S_total, years, "strong", n_devices), "value": S_total * T * len(times) / sum(times),
            "unit": UNIT, "ms_per_step": float(np.mean(times)) * 1e3, "readout_ms": readout * 1e3, "devices": n_devices,
            "failed_scenarios": nfail, "api": "pywr_b200.engine.ShardedEngine: one host process, one engine per GPU, "
            "contiguous global_id blocks, per-scenario recorder results gathered into their column blocks; wall clock"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from benchmarks.workloads import algorithmic_bytes_per_scenario_step, load_workload

    S_arg = args.scenarios or S_DEFAULT.get(args.workload, 1024)
    if args.scaling == "strong" and S_arg % world:
        raise SystemExit(f"--scaling strong: {S_arg} scenarios do not divide over {world} GPUs")
    S = S_arg if args.scaling == "weak" else S_arg // world          # scenarios of this rank
    config = {"workload": workload_text(args.workload, S_arg, args.years, args.scaling, world),
              "scenarios_per_gpu": S, "formulation": "route", "kernel": "register-resident dictionary simplex, "
              "device-resident run, run kernel specialised on the model at load (NVRTC)",
              "l2": "inputs+outputs per step (inflow table + recorder arrays) exceed the 126 MB L2",
              "parallelism": f"scenario-sharded x{world}, no data-path collective"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        s, prog, info = load_workload(args.workload, S, years=args.years)
        config.update(timesteps=info["timesteps"], lp_rows=info["rows"], lp_cols=info["cols"], recorders=info["n_recorders"])
        cores = cpu_cores()
        value, mean_t, nfail = run_cpu(s, prog, info, args.steps, max(args.warmup, 1), cores)
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_t * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"one rank's share of the workload per step ({S} scenarios x {info['timesteps']} timesteps) "
                                       f"on all {cores} host threads; GLPK is not installable in this image, so the CPU arm is "
                                       "the oracle port (oracle/lp_oracle.c, program mode, pthreads over scenario blocks)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "failed_scenarios": nfail,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep rank 0's stdout to the one JSON line (the box's nccl.conf asks for the banner)
        torch.cuda.set_device(local_rank)
        # the communicator's creation prints NCCL's version banner through C stdio on rank 0: fd 1 points at stderr meanwhile
        import ctypes

        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
            ctypes.CDLL(None).fflush(None)
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        cpu_group = dist.new_group(backend="gloo")   # host-side barrier: waiting ranks must not keep a kernel spinning

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

    def measure(workload, S_rank, years, steps, with_e2e, with_full):
        """value / e2e of one configuration, all ranks together: barrier + sync on both sides, max over ranks"""
        b = Bench(workload, S_rank, years, rank * S_rank, local_rank)
        b.warm(max(args.warmup, 3))
        st0 = b.eng.stats()
        barrier()
        t_wall0 = time.perf_counter()
        kernel_ms = b.timed_runs(steps)
        barrier()
        wall = time.perf_counter() - t_wall0
        st1 = b.eng.stats()
        dev_time = max_over_ranks(sum(kernel_ms) / 1e3)
        out = {"bench": b, "kernel_ms": kernel_ms, "wall": wall, "window": (t_wall0, t_wall0 + wall), "st0": st0, "st1": st1,
               "dev_time": dev_time, "value": world * b.units * steps / dev_time, "nfail": b.eng.program_status()[0]}
        for key, full in (("e2e", False), ("e2e_full_arrays", True)):
            if not with_e2e or (full and not with_full):
                out[key] = None
                continue
            step, h2d, d2h = b.e2e_setup(full)
            for _ in range(2):
                step()
            b.eng.sync()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            b.eng.sync()
            barrier()
            e2e_time = max_over_ranks(time.perf_counter() - t0)
            what = ("every [T, S] recorder array and every per-scenario value" if full else
                    "the per-scenario recorder results ([S] accumulators; sum / min / max over time of the array recorders)")
            out[key] = {"value": world * b.units * steps / e2e_time, "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": e2e_time / steps * 1e3,
                        "api": f"b200lp_run_pipelined: reset + upload + run + read-out of {what}, {b.chunks} time slices on "
                               "three streams (pinned host buffers)"}
        return out

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    m = measure(args.workload, S, args.years, args.steps, not args.no_e2e, True)
    b = m["bench"]
    s, prog, info, eng, T, units = b.s, b.prog, b.info, b.eng, b.T, b.units
    config.update(timesteps=T, lp_rows=info["rows"], lp_cols=info["cols"], recorders=info["n_recorders"])
    clocks = sampler.stop(m["window"])
    st0, st1, kernel_ms, value, e2e = m["st0"], m["st1"], m["kernel_ms"], m["value"], m["e2e"]
    launches = int(st1["kernel_launches"] - st0["kernel_launches"])

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
    jit = bool(st1.get("run_kernel_jit"))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_of(args.workload, S, T), "peak_source": peak_src,
                "kernel": "lp_run_jit (run kernel specialised on the model, NVRTC)" if jit else "lp_run_kernel",
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
    b.close()

    # ---- north_star's target as stated: 4,096 thames scenarios in TOTAL over the GPUs of this run ----
    target = None
    if world > 1 and args.scaling == "weak" and args.workload == "thames" and S_arg % world == 0 and not args.no_e2e:
        mt = measure("thames", S_arg // world, args.years, args.steps, True, False)
        target = {"workload": workload_text("thames", S_arg, args.years, "strong", world), "value": mt["value"], "unit": UNIT,
                  "ms_per_step": mt["dev_time"] / args.steps * 1e3, "e2e": mt["e2e"], "failed_scenarios": mt["nfail"]}
        mt["bench"].close()
    # ---- the same target through the product's one-process layer: rank 0 drives all GPUs of the run ----
    single_process = None
    if world > 1 and target is not None:
        import torch

        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                single_process = single_process_run("thames", S_arg, args.years, world, args.steps)
            except Exception as exc:
                single_process = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        dist.barrier(group=cpu_group)
    # ---- BASELINE.json configs[1] beside it (N = 1 only) ---------------------------------------------
    config2 = None
    if world == 1 and args.workload == "thames" and not args.no_e2e:
        m2 = measure("two_reservoir", S_DEFAULT["two_reservoir"], None, args.steps, True, False)
        config2 = {"workload": workload_text("two_reservoir", S_DEFAULT["two_reservoir"], None, "weak", 1), "value": m2["value"],
                   "unit": UNIT, "ms_per_step": m2["dev_time"] / args.steps * 1e3, "e2e": m2["e2e"], "failed_scenarios": m2["nfail"]}
        m2["bench"].close()

    # ---- BASELINE.json configs[3] (the 2,000-node network as SURVEY 8(d) specifies it) on the large-LP path, N = 1 only: its own
    # process (benchmarks/network_program_bench.py), after this process has released its engines
    config4 = None
    if world == 1 and args.workload == "thames" and not args.no_e2e:
        try:
            import subprocess

            cmd = [sys.executable, os.path.join(ROOT, "benchmarks", "network_program_bench.py"), "--fixture", "netspec2000",
                   "--scenarios", "8192", "--steps", "365"]
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
            r4 = json.loads(out.stdout.strip().splitlines()[-1])
            config4 = {"workload": "netspec2000: 2,316-node generated network (150 reservoirs, 40 aggregated nodes, 60 piecewise links), edge "
                                   "formulation, LP 3,105 x 2,010, daily AR(1) inflows per scenario; 8,192 scenarios x 365 days, device-resident",
                       **{k: r4[k] for k in ("value", "unit", "device_s", "wall_s", "reduced_rows", "pivots_per_scenario_timestep", "kernel_variant",
                                             "kernel_launches", "failed_scenarios_max_over_ranks",
                                             "max_rel_diff_vs_reference_host_run_scenarios_0_1", "recorders_checked", "roofline", "roofline_algorithmic")}}
        except Exception as exc:
            config4 = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}

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
            full = years_of(args.workload, args.years)
            drop = {}
            for arrays in ("eager", "lazy"):    # the WHOLE workload through the literal call
                m3 = build_model(args.workload, S, years=full, solver="b200", solver_args={"device_run": True, "device_run_arrays": arrays})
                res3 = m3.run()
                drop[arrays] = {"value": res3.timesteps * S / res3.time_taken, "unit": UNIT, "years": full,
                                "compile_s": res3.solver_stats["compile_s"], "run_s": res3.solver_stats["device_run_s"],
                                "readout_s": res3.solver_stats["readout_s"]}
                del m3, res3
            pywr_extra = {
                "model_run_solver_b200_device_run": dict(
                    drop["eager"], arrays_lazy=drop["lazy"],
                    note="Model.run() of a model whose solver entry says {\"name\": \"b200\", \"device_run\": true}: the literal "
                         "drop-in call, routed to the device-resident run; time = compile + run + read-out into Pywr's objects "
                         "(every [T, S] array by default; arrays_lazy: per-scenario values only, arrays on request)"),
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
            "warmup": max(args.warmup, 3), "ms_per_step": m["dev_time"] / args.steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e, "e2e_full_arrays": m["e2e_full_arrays"], "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu,
            "failed_scenarios": m["nfail"], "wall_ms_per_step_including_reset": m["wall"] / args.steps * 1e3,
            "pivots_per_scenario_timestep": (st1["pivots"] - st0["pivots"]) / max(1, units * args.steps),
            "kernel_variant": st1["kernel_variant"], "run_kernel_jit": jit, "gather_ms": gather_ms,
            "north_star_target": target, "single_process_multi_gpu": single_process, "config2": config2, "config4": config4,
            "through_pywr": pywr_extra,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
