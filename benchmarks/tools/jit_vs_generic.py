"""Run time of a golden program, widened to S scenarios, on the specialised (NVRTC) and on the generic run kernel.
usage: python benchmarks/tools/jit_vs_generic.py <golden program case> <scenarios> [repeat-years]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fixtures_util import load_program_case, widen_program  # noqa: E402
from pywr_b200.engine import CudaEngine  # noqa: E402


def main():
    name, S = sys.argv[1], int(sys.argv[2])
    s, prog, _, _ = load_program_case(name)
    q = widen_program(s, prog, S, seed=1)
    out = {}
    for jit in ("1", "0"):
        os.environ["B200LP_JIT"] = jit
        eng = CudaEngine(s, S, device=0)
        eng.load_program(q)
        best = 1e9
        for _ in range(4):
            eng.program_reset()
            eng.mark(0); eng.run(0, q.n_steps); eng.mark(1)
            best = min(best, eng.elapsed_ms())
        st = eng.stats()
        out[jit] = (best, st["run_kernel_jit"], st["pivots"])
        eng.close()
    units = S * q.n_steps
    print(f"{name} S={S} T={q.n_steps}: specialised {out['1'][0]:.3f} ms ({units / out['1'][0] / 1e3:.0f} M/s, jit={out['1'][1]}), "
          f"generic {out['0'][0]:.3f} ms ({units / out['0'][0] / 1e3:.0f} M/s), speed-up {out['0'][0] / out['1'][0]:.2f}")


if __name__ == "__main__":
    main()
