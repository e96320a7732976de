"""Independent pin of the config-4 fixture (BUILD container; a few minutes): the committed device program of the 2,073-node
network (benchmarks/fixtures/network600.program.npz, 365 daily timesteps, 2 scenarios) executed by the DICTIONARY oracle's
program mode (oracle/lp_oracle.c: dense bounded dual simplex, an implementation independent of the revised simplex in
pywr_b200/csrc/rsx_core.cuh whose host build produced network600.expected.npz).  Writes network600.expected_dict.npz.

usage: python tests/golden/make_network_program_dict.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.oracle_engine import OracleEngine  # noqa: E402
from pywr_b200.program import Program  # noqa: E402
from pywr_b200.structure import LPStructure  # noqa: E402

FIX = os.path.join(ROOT, "benchmarks", "fixtures", "network600")


def main():
    s = LPStructure.load(FIX + ".pstructure.npz")
    prog = Program.load(FIX + ".program.npz")
    S = int(prog.initial_volume.size // s.n_storage)
    eng = OracleEngine(s, S, threads=S)
    eng.load_program(prog)
    t0 = time.time()
    eng.run(0, prog.n_steps)
    n_failed = eng.program_status()[0]
    assert n_failed == 0
    outs = {f"rec{r}": eng.fetch_recorder(r, prog.rec_shape(r, S)) for r in range(prog.n_recorders)}
    c = eng.counters()
    np.savez_compressed(FIX + ".expected_dict.npz", names=np.array(prog.rec_names), pivots=np.array(c["pivots"]), **outs)
    print(f"network600 on the dictionary oracle: S={S} T={prog.n_steps} {c['pivots']} pivots, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
