"""Compiles a dumped specialised run kernel (B200LP_JIT_DUMP=<file>) with NVRTC exactly like csrc/jit_run.inc does and
writes <file>.cubin (for cuobjdump -sass / nvdisasm -g).  With an ncu source-page CSV (ncu -i rep --page source
--print-source sass --csv) it prints executed instructions per source line of the dumped kernel.
usage: python benchmarks/tools/jit_sass.py kernel.cu [source_page.csv scenario_timesteps]"""
import collections
import csv
import ctypes
import re
import subprocess
import sys


def compile_cubin(path):
    nv = ctypes.CDLL("libnvrtc.so.12")
    src = open(path, "rb").read()
    prog = ctypes.c_void_p()
    assert nv.nvrtcCreateProgram(ctypes.byref(prog), src, b"lp_run_jit.cu", 0, None, None) == 0
    opts = [b"--gpu-architecture=sm_100a", b"--fmad=false", b"--std=c++17", b"-lineinfo", b"--ptxas-options=-v", b"-default-device"]
    arr = (ctypes.c_char_p * len(opts))(*opts)
    rc = nv.nvrtcCompileProgram(prog, len(opts), arr)
    n = ctypes.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, ctypes.byref(n))
    log = ctypes.create_string_buffer(n.value)
    nv.nvrtcGetProgramLog(prog, log)
    assert rc == 0, log.value.decode()
    nv.nvrtcGetCUBINSize(prog, ctypes.byref(n))
    cub = ctypes.create_string_buffer(n.value)
    nv.nvrtcGetCUBIN(prog, cub)
    open(path + ".cubin", "wb").write(cub.raw)
    return log.value.decode()


def line_map(cubin):
    """offset -> (source line, sass text) of lp_run_jit"""
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    cur, infn, m = None, False, {}
    for l in out:
        if ".text." in l:
            infn = ".text.lp_run_jit" in l
        g = re.search(r'//## File ".*", line (\d+)', l)
        if g:
            cur = int(g.group(1))
            continue
        if not infn:
            continue
        g = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if g:
            m[int(g.group(1), 16)] = (cur, g.group(2))
    return m


def main():
    path = sys.argv[1]
    log = compile_cubin(path)
    print([l for l in log.splitlines() if "registers" in l][-1])
    if len(sys.argv) < 4:
        return
    units = float(sys.argv[3])
    lm = line_map(path + ".cubin")
    src = open(path).read().splitlines()
    rows = list(csv.reader(open(sys.argv[2])))
    hdr = rows[1]
    iA, iN, iI = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    base, per_line, samples = None, collections.Counter(), collections.Counter()
    for r in rows[2:]:
        try:
            a, n, ins = int(r[iA], 16), int(r[iN]), int(r[iI])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = a
        ln = lm.get(a - base, (None, ""))[0]
        per_line[ln] += ins
        samples[ln] += n
    tot, tots = sum(per_line.values()), sum(samples.values())
    print(f"total {tot / units:.1f} warp instructions per scenario-timestep, {tots} samples")
    for ln, ins in sorted(per_line.items(), key=lambda kv: -kv[1])[:60]:
        text = src[ln - 1].strip()[:110] if ln else "?"
        print(f"{ins / units:7.1f} inst  {100.0 * samples[ln] / tots:5.1f} % samples  L{ln}: {text}")


if __name__ == "__main__":
    main()
