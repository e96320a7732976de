"""CPU tests of the one-process multi-GPU layer (pywr_b200.engine.ShardedEngine, Program.shard): the scenario batch
cut into contiguous global_id blocks (pywr/_core.pyx:210-240) gives bit-identical results to the whole batch.  The
shards here are oracle engines (no GPU in this container); tests/test_gpu_program.py repeats it on CUDA engines."""
import numpy as np
import pytest

from fixtures_util import load_program_case, run_program, widen_program
from oracle.oracle_engine import OracleEngine
from pywr_b200.engine import ShardedEngine, scenario_blocks


def test_scenario_blocks_are_contiguous_and_cover_the_batch():
    assert scenario_blocks(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert scenario_blocks(2, 8) == [(0, 1), (1, 2)]
    assert scenario_blocks(4096, 8) == [(512 * k, 512 * (k + 1)) for k in range(8)]


@pytest.mark.parametrize("name,S,shards", [("two_reservoir.route", 10, 3), ("thames.route", 7, 4), ("demand_saving2.route", 5, 2),
                                           ("rolling_virtual_storage_long.route", 6, 3), ("lagged_flows.route", 9, 4),
                                           ("derived_parameters.route", 8, 3)])
def test_sharded_engine_equals_single_engine(name, S, shards):
    s, prog, _, names = load_program_case(name)
    q = widen_program(s, prog, S, seed=17)
    one = OracleEngine(s, S)
    many = ShardedEngine(s, S, devices=shards, factory=lambda st, n, d: OracleEngine(st, n, threads=1))
    assert [e.S for e in many.shards] == [hi - lo for lo, hi in scenario_blocks(S, shards)]
    o1, st1, x1 = run_program(one, q)
    o2, st2, x2 = run_program(many, q, chunks=[5, 40])
    assert st1 == st2
    for a, b, nm in zip(o1, o2, names):
        assert np.array_equal(a, b, equal_nan=True), nm
    for a, b in zip(x1, x2):
        assert np.array_equal(a, b, equal_nan=True)
    assert many.counters()["pivots"] == one.counters()["pivots"]
    # temporal aggregates and table updates go to the right column blocks
    for r in range(q.n_recorders):
        if q.rec_shape(r, S) != (S,):
            for a, b in zip(one.fetch_recorder_agg(r), many.fetch_recorder_agg(r)):
                assert np.array_equal(a, b, equal_nan=True)
    per_scenario = np.nonzero(q.table_cols == S)[0]
    if per_scenario.size == 0:                 # no time series in this model (nothing was widened)
        one.close(); many.close()
        return
    k = int(per_scenario[0])
    r0 = int(q.table_rows[k])
    new = q.table_data[q.table_offset[k]: q.table_offset[k] + r0 * S].reshape(r0, S) * 0.5
    for e in (one, many):
        e.update_table(k, np.ascontiguousarray(new))
        e.program_reset()
        e.run(0, q.n_steps)
        e.sync()
    for r in range(q.n_recorders):
        assert np.array_equal(one.fetch_recorder(r, q.rec_shape(r, S)), many.fetch_recorder(r, q.rec_shape(r, S)), equal_nan=True)
    one.close(); many.close()


def test_failed_scenario_is_reported_with_its_global_index():
    s, prog, _, _ = load_program_case("thames.route")
    S = 6
    q = widen_program(s, prog, S, seed=3)
    k = int(np.nonzero(q.table_cols == S)[0][0])
    r0 = int(q.table_rows[k])
    data = q.table_data[q.table_offset[k]: q.table_offset[k] + r0 * S].reshape(r0, S).copy()
    data[3, 4] = np.nan                       # scenario 4 fails at timestep 3
    q.table_data = q.table_data.copy()
    q.table_data[q.table_offset[k]: q.table_offset[k] + r0 * S] = data.reshape(-1)
    one = OracleEngine(s, S)
    many = ShardedEngine(s, S, devices=3, factory=lambda st, n, d: OracleEngine(st, n, threads=1))
    _, st1, _ = run_program(one, q)
    _, st2, _ = run_program(many, q)
    assert st1 == st2 and st1[0] == 1 and st1[1] == 4 and st1[3] == 3
    one.close(); many.close()
