"""Turns an ncu report (gpurun_out/*.ncu-rep, scratch) into the tracked summary under profiles/.
usage: python profiles/summarize.py <report.ncu-rep> <scenario-timesteps in the profiled launch> <out.md> [title]"""
import collections
import csv
import re
import subprocess
import sys


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, units, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    title = sys.argv[4] if len(sys.argv) > 4 else rep
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, unit, vals = rows[0], rows[1], rows[2]
    get = {h: (vals[i], unit[i]) for i, h in enumerate(hdr)}
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
            "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_average_branch_targets_threads_uniform.pct",
            "sm__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_shared_st.sum"]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on)", "",
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
    src = list(csv.reader(ncu(rep, "--page", "source", "--print-source", "cuda,sass", "--csv").splitlines()))
    h2 = src[2]
    iS, iI = h2.index("# Samples"), h2.index("Instructions Executed")
    agg, ops, tot_s = collections.OrderedDict(), collections.Counter(), 0
    for r in src[3:]:
        try:
            s, ins = int(r[iS]), int(r[iI])
        except (ValueError, IndexError):
            continue
        if r[0] != "":
            a = agg.setdefault((int(r[0]), r[1].strip()[:110]), [0, 0])
            a[0] += s; a[1] += ins; tot_s += s
        else:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
            if m:
                ops[m.group(2)] += ins
    lines += ["", "## hottest source lines (pywr_b200/csrc/lp_kernel.cuh)", "", "| line | instr / scenario-timestep | stall samples % | source |",
              "|---|---|---|---|"]
    for (ln, text), (s, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        lines.append(f"| {ln} | {ins / units:.1f} | {100 * s / max(tot_s, 1):.1f} | `{text.replace('|', '/')}` |")
    tot_i = sum(ops.values())
    lines += ["", "## SASS opcode mix", "", "| opcode | per scenario-timestep | share % |", "|---|---|---|"]
    for op, n in ops.most_common(16):
        lines.append(f"| {op} | {n / units:.1f} | {100 * n / max(tot_i, 1):.1f} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
