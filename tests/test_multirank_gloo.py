"""N > 1 on CPU: two ranks (gloo), each running its own contiguous block of scenarios (the sharding
bench.py uses: no data-path collective), gather the per-scenario recorder values and compare with a
single-rank run over the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S_local, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from benchmarks.workloads import load_workload
    from fixtures_util import run_program
    from oracle.oracle_engine import OracleEngine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, prog, _ = load_workload("two_reservoir", S_local, years=1, first_scenario=rank * S_local)
    eng = OracleEngine(s, S_local)
    outs, status, _ = run_program(eng, prog)
    assert status[0] == 0
    gathered = []
    for r, o in enumerate(outs):
        if o.ndim == 1:   # per-scenario values: gather in rank order = global scenario order
            t = torch.from_numpy(np.ascontiguousarray(o))
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            gathered.append(torch.cat(parts).numpy())
    total = torch.tensor([float(np.sum(outs[4]))], dtype=torch.float64)   # e.g. total transferred volume
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), *gathered, total=total.numpy())
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


def test_two_ranks_equal_one_rank(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from benchmarks.workloads import load_workload
    from fixtures_util import run_program
    from oracle.oracle_engine import OracleEngine

    world, S_local = 2, 5
    port = _free_port()
    mp.spawn(_worker, args=(world, port, S_local, str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "gathered.npz")
    s, prog, _ = load_workload("two_reservoir", world * S_local, years=1)
    eng = OracleEngine(s, world * S_local)
    outs, status, _ = run_program(eng, prog)
    ref = [o for o in outs if o.ndim == 1]
    assert len(ref) == len([k for k in z.files if k.startswith("arr_")])
    for i, o in enumerate(ref):
        assert np.array_equal(z[f"arr_{i}"], o)      # scenario indexing is bit-exact across the split
    assert np.isclose(z["total"][0], np.sum(outs[4]), rtol=1e-13)
    eng.close()
