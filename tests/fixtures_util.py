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


def synth_batch(structure, trace, S, T, seed):
    """A scenario batch of size S built from a recorded single-model trace: every recorded plane is
    tiled over scenarios and perturbed (flows scaled log-normally, volumes redrawn uniformly within
    the storage bounds), so that scenarios take different paths through the simplex."""
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
