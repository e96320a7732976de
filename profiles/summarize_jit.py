"""Turns an ncu report of the specialised run kernel (lp_run_jit, compiled by NVRTC from generated source) into a tracked
summary.  The source lines come from the dump of the generated kernel (B200LP_JIT_DUMP=<file> of the same run).
usage: python profiles/summarize_jit.py <report.ncu-rep> <dumped kernel.cu> <scenario-timesteps in the launch> <out.md> [title]"""
import collections
import csv
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "benchmarks", "tools"))
import jit_sass  # noqa: E402


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, dump, units, out = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    title = sys.argv[5] if len(sys.argv) > 5 else rep
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, unit, vals = rows[0], rows[1], rows[2]
    get = {h: (vals[i], unit[i]) for i, h in enumerate(hdr)}
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
            "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__sass_inst_executed_op_shared_ld.sum",
            "sm__sass_inst_executed_op_shared_st.sum", "smsp__sass_average_branch_targets_threads_uniform.pct"]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on), generated kernel `{dump}`", "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in keys:
        if k in get:
            lines.append(f"| {k} | {get[k][0]} | {get[k][1]} |")
    inst = float(get["smsp__inst_executed.sum"][0])
    lines += ["", f"scenario-timesteps in this launch: {units:.0f} -> **{inst / units:.1f} warp instructions per scenario-timestep**", "",
              "## warp stall reasons (cycles per issued instruction)", "", "| reason | value |", "|---|---|"]
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                v = float(vals[i])
            except ValueError:
                continue
            if v > 0.02:
                lines.append(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |")
    jit_sass.compile_cubin(dump)
    lm = jit_sass.line_map(dump + ".cubin")
    src_text = open(dump).read().splitlines()
    src = list(csv.reader(ncu(rep, "--page", "source", "--print-source", "sass", "--csv").splitlines()))
    h2 = src[1]
    iA, iS, iI, iT = h2.index("Address"), h2.index("# Samples"), h2.index("Instructions Executed"), h2.index("Source")
    per_line, samples, ops, hot, base = collections.Counter(), collections.Counter(), collections.Counter(), 0.0, None
    for r in src[2:]:
        try:
            a, s, ins = int(r[iA], 16), int(r[iS]), int(r[iI])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = a
        ln = lm.get(a - base, (None, ""))[0]
        per_line[ln] += ins
        samples[ln] += s
        if ins / units > 0.5:
            hot += ins / units
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iT])
        if m:
            ops[m.group(2)] += ins
    tot_s = sum(samples.values())
    lines += ["", f"instructions executed on (almost) every timestep: **{hot:.0f} per scenario-timestep** (the steady-state path); the rest "
              "is the general simplex of the timesteps that pivot", "",
              "## hottest source lines (line numbers of the generated kernel; lines below the kernel's first line are lp_kernel.cuh)", "",
              "| line | instr / scenario-timestep | stall samples % | source |", "|---|---|---|---|"]
    for ln, ins in sorted(per_line.items(), key=lambda kv: -samples[kv[0]])[:30]:
        text = src_text[ln - 1].strip()[:110] if ln else "?"
        lines.append(f"| {ln} | {ins / units:.1f} | {100 * samples[ln] / max(tot_s, 1):.1f} | `{text.replace('|', '/')}` |")
    tot_i = sum(ops.values())
    lines += ["", "## SASS opcode mix", "", "| opcode | per scenario-timestep | share % |", "|---|---|---|"]
    for op, n in ops.most_common(16):
        lines.append(f"| {op} | {n / units:.1f} | {100 * n / max(tot_i, 1):.1f} |")
    open(out, "w").write("\n".join(lines) + "\n")
    os.remove(dump + ".cubin")
    print("wrote", out)


if __name__ == "__main__":
    main()
