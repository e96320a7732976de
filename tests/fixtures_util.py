"""Helpers shared by the tests: golden fixture loading, trace replay, synthetic scenario batches."""
import glob
import os

import numpy as np

from pywr_b200.structure import LPStructure

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[: -len(".trace.npz")] for p in glob.glob(os.path.join(GOLDEN, "*.trace.npz")))


def load_case(name):
    s = LPStructure.load(os.path.join(GOLDEN, name + ".structure.npz"))
    z = np.load(os.path.join(GOLDEN, name + ".trace.npz"))
    return s, {k: z[k] for k in z.files}


def replay(engine, structure, slots_seq, days_seq, want_cols=True):
    """Feeds a recorded sequence of input planes through an engine (oracle or CUDA, same duck type).
    Returns node_flow [T,N,S], col_flow [T,n,S], objective [T,S], nbad [T]."""
    T = len(days_seq)
    S = engine.S
    slots = engine.alloc_planes(max(structure.n_slots, 1))
    nf = engine.alloc_planes(structure.n_nodes)
    cf = engine.alloc_planes(structure.n_cols)
    ob = engine.alloc_planes(1)
    out_nf = np.zeros((T, structure.n_nodes, S))
    out_cf = np.zeros((T, structure.n_cols, S))
    out_ob = np.zeros((T, S))
    nbad = np.zeros(T, dtype=np.int64)
    engine.reset(None)
    for t in range(T):
        if structure.n_slots:
            slots[: structure.n_slots] = slots_seq[t]
        nbad[t] = engine.step(float(days_seq[t]), slots, nf, cf if want_cols else None, ob)
        out_nf[t] = nf
        out_cf[t] = cf
        out_ob[t] = ob[0]
    return out_nf, out_cf, out_ob, nbad


def synth_batch(structure, trace, S, T, seed, common=False):
    """A scenario batch of size S built from a recorded single-model trace: every recorded plane is
    tiled over scenarios and perturbed (flows scaled log-normally, volumes redrawn uniformly within
    the storage bounds), so that scenarios take different paths through the simplex.  `common`: one
    factor per scenario for all planes (a wetter / drier scenario) - keeps the edge-form LPs, whose
    planes are coupled through mass-balance rows, feasible."""
    rng = np.random.default_rng(seed)
    base = trace["slots"]  # [T0, n_slots, S0]
    T0, n_slots, S0 = base.shape
    T = min(T, T0)
    out = np.empty((T, n_slots, S))
    pick = rng.integers(0, S0, size=S)
    vol_slots = {int(r) for r in structure.st_vol_ref if r >= 0}
    flag_slots = {int(r) for r in structure.st_active_ref if r >= 0}
    maxv_of = {int(v): (int(mx), int(mn)) for v, mx, mn in zip(structure.st_vol_ref, structure.st_maxv_ref,
                                                                 structure.st_minv_ref) if v >= 0}
    scale = np.exp(rng.normal(0.0, 0.4, size=(n_slots, S)))
    if common:
        scale = np.repeat(np.exp(rng.normal(0.0, 0.4, size=(1, S))), n_slots, axis=0)
    for t in range(T):
        plane = base[t][:, pick]
        for k in range(n_slots):
            if k in vol_slots or k in flag_slots:
                out[t, k] = plane[k]
            else:
                v = plane[k] * scale[k]
                out[t, k] = np.where(np.isfinite(plane[k]), v, plane[k])
    # volumes: uniform in [min, max] (the engine treats the volume as an input in this mode)
    for k, (mx, mn) in maxv_of.items():
        for t in range(T):
            hi = out[t, mx] if mx >= 0 else structure.consts[-mx - 1]
            lo = out[t, mn] if mn >= 0 else structure.consts[-mn - 1]
            hi = np.broadcast_to(hi, (S,)); lo = np.broadcast_to(lo, (S,))
            u = rng.random(S)
            # a few exactly full / exactly empty reservoirs to hit the FX / clipped-bound branches
            u[rng.random(S) < 0.05] = 0.0
            u[rng.random(S) < 0.05] = 1.0
            out[t, k] = lo + u * (hi - lo)
    return out, trace["days"][:T]


# ---- device-run ("program") fixtures -------------------------------------------------------------
def program_cases():
    return sorted(os.path.basename(p)[: -len(".program.npz")] for p in glob.glob(os.path.join(GOLDEN, "*.program.npz")))


def load_program_case(name):
    from pywr_b200.program import Program

    s = LPStructure.load(os.path.join(GOLDEN, name + ".pstructure.npz"))
    prog = Program.load(os.path.join(GOLDEN, name + ".program.npz"))
    z = np.load(os.path.join(GOLDEN, name + ".expected.npz"))
    names = [str(n) for n in z["names"]]      # the model's own recorders; hidden ones (lag state) follow them in the program
    expected = [z[f"rec{i}"] for i in range(len(names))]
    return s, prog, expected, names


def n_scenarios_of(s, prog):
    return int(prog.initial_volume.size // max(s.n_storage, 1)) if s.n_storage else int(prog.scen_index.size // max(prog.n_axes, 1))


def run_program(engine, prog, chunks=None):
    """Executes a loaded program on an engine (oracle or CUDA) and returns all recorder outputs."""
    engine.load_program(prog)
    T = prog.n_steps
    if chunks is None:
        engine.run(0, T)
    else:
        t = 0
        for c in chunks:
            c = min(c, T - t)
            if c <= 0:
                break
            engine.run(t, c)
            t += c
        if t < T:
            engine.run(t, T - t)
    engine.sync()
    status = engine.program_status()
    outs = [engine.fetch_recorder(r, prog.rec_shape(r, engine.S)) for r in range(prog.n_recorders)]
    return outs, status, engine.fetch_state()


def widen_program(s, prog, S, seed, sigma=0.35):
    """Scenario batch of size S from a single-scenario program: every time-varying single-column table
    becomes a per-scenario table (column = scenario id) whose columns are the original series scaled
    by a smooth log-normal AR(1) factor, so that scenarios diverge (different bases, different
    control-curve crossings).  Initial state is tiled."""
    import copy

    from pywr_b200.program import COL_SCENARIO

    S0 = n_scenarios_of(s, prog)
    rng = np.random.default_rng(seed)
    q = copy.copy(prog)
    T = prog.n_steps
    rows, cols, axis, offs = list(prog.table_rows), list(prog.table_cols), list(prog.table_axis), []
    chunks, pos = [], 0
    pick = rng.integers(0, S0, size=S)
    for t in range(prog.n_tables):
        data = prog.table_data[prog.table_offset[t]: prog.table_offset[t] + rows[t] * cols[t]].reshape(rows[t], cols[t])
        if rows[t] == T and (cols[t] == 1 or axis[t] >= 0) and np.ptp(data) > 0:
            if axis[t] >= 0:
                base = data[:, prog.scen_index.reshape(prog.n_axes, S0)[axis[t]][pick]]
            else:
                base = np.repeat(data, S, axis=1)
            z = np.zeros((T, S))
            e = rng.normal(0.0, sigma * np.sqrt(1 - 0.98**2), size=(T, S))
            z[0] = rng.normal(0.0, sigma, size=S)
            for i in range(1, T):
                z[i] = 0.98 * z[i - 1] + e[i]
            data = base * np.exp(z)
            cols[t], axis[t] = S, COL_SCENARIO
        offs.append(pos)
        chunks.append(np.ascontiguousarray(data).reshape(-1))
        pos += data.size
    q.table_rows, q.table_cols = np.array(rows, np.int32), np.array(cols, np.int32)
    q.table_axis, q.table_offset = np.array(axis, np.int32), np.array(offs, np.int64)
    q.table_data = np.concatenate(chunks) if chunks else np.zeros(0)
    q.scen_index = prog.scen_index.reshape(max(prog.n_axes, 1), S0)[:, pick].reshape(-1).astype(np.int32)
    nst = max(s.n_storage, 1)
    q.initial_volume = prog.initial_volume.reshape(nst, S0)[:, pick].reshape(-1)
    q.initial_pc = prog.initial_pc.reshape(nst, S0)[:, pick].reshape(-1)
    return q


def rel_diff(a, b):
    """max |a-b| / max(1,|b|), treating identical infinities / NaNs as equal"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    d = np.where(same, 0.0, d)
    return float(np.max(np.where(np.isnan(d), np.inf, d)))
