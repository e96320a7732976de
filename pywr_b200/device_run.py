"""Run a whole Pywr model on the device and hand the results back to Pywr's own objects.

``run_on_device(model)`` is the device-resident counterpart of ``Model.run()``
(``pywr/_model.pyx:637-682``): it performs the same setup / reset, executes every timestep for every
scenario inside the engine (no per-timestep host work), then writes the recorder arrays, final node
flows and storage volumes into the reference's objects through the Cython bridge and calls
``Model.finish()``.  ``recorder.data``, ``values()``, ``aggregated_value()`` and ``to_dataframe()``
then work unchanged.
"""

from __future__ import annotations

import time

import numpy as np

from .program import ARRAY_KINDS, PROG_AGG_ONLY, Unsupported, _flow_form, compile_program

# temporal aggregation functions the device accumulates while it writes an array recorder (b200lp_fetch_recorder_agg)
DEVICE_TEMPORAL_AGGS = ("sum", "mean", "min", "max")


class DeviceRunResult(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def close(self):
        """Releases the engine a run with ``arrays="lazy"`` keeps for later array fetches."""
        rec = self.get("recorders")
        if isinstance(rec, LazyRecorders):
            rec.close()


class LazyRecorders(dict):
    """Recorder results of a run with ``arrays="lazy"``: per-scenario values are here already; the [T, S] array of an
    array recorder is copied from the device when :meth:`data` asks for it (and cached)."""

    def __init__(self, eng, prog, S):
        super().__init__()
        self._eng, self._prog, self._S, self._arrays = eng, prog, S, {}
        self._index = {rec.name: r for r, rec in enumerate(prog.recorders)}

    def data(self, name):
        r = self._index[name]
        if int(self._prog.rec_kind[r]) not in ARRAY_KINDS:
            return self[name]
        if name not in self._arrays:
            if self._eng is None:
                raise RuntimeError("the device run's engine has been closed")
            self._arrays[name] = self._eng.fetch_recorder(r, self._prog.rec_shape(r, self._S))
        return self._arrays[name]

    def close(self):
        if self._eng is not None:
            self._eng.close()
            self._eng = None


def park_timestepper(model):
    """Leaves the timestepper where a host run leaves it — ``current`` one past the end (pywr/timestepper.py:86-103),
    ``model.timestep`` the last timestep — without stepping through the whole calendar: MeanFlow / DeficitFrequency
    recorders divide by ``timestepper.current.index`` in ``finish()``."""
    from pywr import _core

    ts = model.timestepper
    ts.reset()
    last = len(ts._periods) - 1
    ts._next = _core.Timestep(ts._periods[last], last, ts._deltas[last])
    model.timestep = ts.next()
    try:
        ts.next()
    except StopIteration:
        pass


PLAIN_ARRAY_RECORDERS = {"NumpyArrayNodeRecorder", "NumpyArrayNodeDeficitRecorder", "NumpyArrayNodeSuppliedRatioRecorder",
                         "NumpyArrayNodeCurtailmentRatioRecorder", "NumpyArrayStorageRecorder", "NumpyArrayParameterRecorder",
                         "NumpyArrayLevelRecorder", "NumpyArrayAreaRecorder"}


def temporal_aggregate(rec, T, ssum, smin, smax):
    """Recorder.values() of an array recorder from the device's running aggregates, or None when its temporal
    aggregation needs the array itself (median, percentiles, custom callables, ignore_nan)."""
    func = device_temporal_agg(rec)
    if func is None:
        return None
    return {"sum": ssum, "mean": ssum / T, "min": smin, "max": smax}[func]


def device_temporal_agg(rec):
    """"sum" / "mean" / "min" / "max" when the device's running aggregates give this recorder's values(), else None."""
    from .bridge import _bridge

    if type(rec).__name__ not in PLAIN_ARRAY_RECORDERS:      # duration curves etc.: values() is not a temporal aggregate of the array
        return None
    func = _bridge.temporal_agg_func(rec)
    if not isinstance(func, str) or func.lower() not in DEVICE_TEMPORAL_AGGS or getattr(rec, "ignore_nan", False):
        return None
    return func.lower()


def run_on_device(model, engine_factory=None, formulation="route", chunk=None, writeback=True, threads=None,
                  host_mirror=True, devices=None, arrays="eager"):
    """Returns a :class:`DeviceRunResult` (timings, counters, speed in scenario-timesteps/s).

    ``engine_factory(structure, n_scenarios)`` defaults to the CUDA engine.  Raises
    :class:`pywr_b200.program.Unsupported` if the model needs host-side evaluation (use
    ``Model.run()`` with ``solver="b200"`` then), and ``RuntimeError`` if a scenario's LP fails, as
    the reference does (``cython_glpk.pyx:1044-1061``).

    Host mirror protocol (``host_mirror``): recorders the device cannot evaluate
    (parameter / event / custom Python recorders ...) stay on the host.  Only when such recorders exist
    the run proceeds one timestep per launch and, after each, the node flows and storage volumes are
    copied back into Pywr's objects (``_flow``, ``_volume``, ``_current_pc``) so that
    ``Model.before()`` (host copies of the parameters) and the recorders' ``after()`` see exactly the
    state of a host run; models without them pay nothing.

    ``devices``: ``None`` = one GPU (``LOCAL_RANK``); ``"all"``, a count or a list of CUDA device indices = the scenario
    batch is cut into contiguous ``global_id`` blocks, one per GPU, all driven by this process
    (:class:`pywr_b200.engine.ShardedEngine`); results are identical to the one-GPU run.

    ``arrays``: what happens to the ``[T, S]`` arrays of the NumpyArray recorders —
    ``"eager"`` (default) copies them into the recorder objects like a host run;
    ``"lazy"`` leaves them on the device: the recorder objects receive ONE row, their temporal aggregate per scenario
    (accumulated on the device), so ``values()`` / ``aggregated_value()`` are unchanged, and ``result.recorders.data(name)``
    fetches an array when it is wanted (call ``result.close()`` afterwards);
    ``"none"`` does not even store them (``B200LP_PROG_AGG_ONLY``): objectives of optimisation batches and ensembles.
    With ``"lazy"`` / ``"none"`` an array recorder whose temporal aggregation is not sum / mean / min / max (or that
    ignores NaNs) is fetched eagerly / raises :class:`Unsupported`."""
    from .bridge import _bridge

    t0 = time.perf_counter()
    if model.dirty or model.timestepper.dirty:
        model.setup()
    else:
        model.reset()
    t1 = time.perf_counter()
    if arrays not in ("eager", "lazy", "none"):
        raise ValueError("arrays must be 'eager', 'lazy' or 'none'")
    s, prog = compile_program(model, formulation, allow_host_recorders=host_mirror)
    S, T = len(model.scenarios.combinations), prog.n_steps
    host_recs = list(prog.host_recorders)
    if arrays == "none":
        for r, rec in enumerate(prog.recorders):
            if int(prog.rec_kind[r]) in ARRAY_KINDS and device_temporal_agg(rec) is None:
                raise Unsupported(f"arrays='none': recorder {rec.name!r} aggregates over time with "
                                  f"{_bridge.temporal_agg_func(rec)!r}" + (" ignoring NaNs" if rec.ignore_nan else "") +
                                  ", which needs the array")
        if not prog.has_lag:     # lagged flows read yesterday's sample from the recorder's array: it has to exist
            prog.flags |= PROG_AGG_ONLY
    if engine_factory is None:
        from .engine import CudaEngine, ShardedEngine

        def engine_factory(structure, n_scenarios):
            if devices is not None:
                return ShardedEngine(structure, n_scenarios, devices=devices, for_program=True)
            return CudaEngine(structure, n_scenarios, for_program=True)
    eng = engine_factory(s, S)
    keep_engine = False
    try:
        eng.load_program(prog)
        t2 = time.perf_counter()
        if host_recs:
            _run_mirrored(model, eng, s, prog, host_recs, _bridge)
        else:
            step = chunk or T
            for a in range(0, T, step):
                eng.run(a, min(step, T - a))
        eng.sync()
        t3 = time.perf_counter()
        n_failed, scen, code, tstep = eng.program_status()
        if n_failed:
            from .solver import B200LPError, _STATUS_TEXT

            if code == 4:
                raise B200LPError(f"NaN detected in the LP data of scenario {scen} at timestep {tstep}")
            raise RuntimeError('Simplex solver failed with message: "{}", status: "{}" '
                               "(scenario {}, timestep {}; {} scenario(s) failed).".format(
                                   _STATUS_TEXT.get(code, "solver failed"), "solution is undefined", scen, tstep, n_failed))
        results = {} if arrays != "lazy" else LazyRecorders(eng, prog, S)
        keep_engine = arrays == "lazy"
        for r, rec in enumerate(prog.recorders):
            is_array = int(prog.rec_kind[r]) in ARRAY_KINDS
            if is_array and arrays != "eager":
                row = temporal_aggregate(rec, T, *eng.fetch_recorder_agg(r))
                if row is not None:
                    results[rec.name] = row
                    if writeback:
                        _bridge.replace_array_recorder(rec, np.ascontiguousarray(row[None, :]))
                    continue
            out = eng.fetch_recorder(r, prog.rec_shape(r, S))
            results[rec.name] = out
            if writeback:
                if is_array:
                    try:        # the fetched [T, S] array BECOMES the recorder's array: no second pass over it
                        _bridge.replace_array_recorder(rec, out)
                    except TypeError:
                        _bridge.set_array_recorder(rec, out)
                else:
                    _bridge.set_constant_recorder(rec, out)
        if writeback:
            x, vol, pc = eng.fetch_state()
            node_id = {id(n): i for i, n in enumerate(s.all_nodes)}
            for node in s.all_nodes:
                try:
                    form = _flow_form(s, node, node_id)
                except Unsupported:
                    continue
                flow = np.zeros(S)
                for c in sorted(form):
                    flow += form[c] * x[c]
                _bridge.set_node_flow(node, flow)
            for k, st in enumerate(s.storages):
                _bridge.set_storage_state(st, vol[k], pc[k])
            # park the timestepper one past the end, as after a host run (pywr/timestepper.py:86-103):
            # MeanFlow / DeficitFrequency recorders divide by current.index in finish()
            if not host_recs:
                park_timestepper(model)
            model.finish()
        t4 = time.perf_counter()
        counters = eng.counters()
        try:
            counters = dict(counters, kernel_variant=eng.stats().get("kernel_variant"))
        except Exception:   # engines without stats() (test doubles)
            pass
    except BaseException:
        keep_engine = False
        raise
    finally:
        if not keep_engine:
            eng.close()
    return DeviceRunResult(
        num_scenarios=S, timesteps=T, time_setup=t1 - t0, time_compile=t2 - t1, time_run=t3 - t2,
        time_readout=t4 - t3, speed=S * T / (t3 - t2) if t3 > t2 else float("nan"), recorders=results,
        counters=counters, structure=s, program=prog, host_recorders=[getattr(r, "name", None) for r in host_recs],
    )


def _node_flow_writer(s, _bridge):
    """(node, sorted column list, weights) for every node whose flow is a linear form of the columns."""
    node_id = {id(n): i for i, n in enumerate(s.all_nodes)}
    out = []
    for node in s.all_nodes:
        try:
            form = _flow_form(s, node, node_id)
        except Unsupported:
            continue
        cols = sorted(form)
        out.append((node, np.asarray(cols, np.int64), np.asarray([form[c] for c in cols])))
    return out


def _run_mirrored(model, eng, s, prog, host_recs, _bridge):
    """One timestep per launch with the host mirrors refreshed in between (host mirror protocol).

    Order of one timestep = ``Model._step`` (``pywr/_model.pyx:629-635``): ``before()`` on the host (node
    ``before`` + parameter values from the mirrored state of the previous timestep), solve + commit +
    ``Storage.after`` on the device, mirrors of ``_flow`` / ``_volume`` / ``_current_pc`` refreshed, then
    ``after()`` of the host-side components (parameters and the host recorders); the device recorders
    accumulate on the device as usual."""
    from pywr.parameters import Parameter

    writers = _node_flow_writer(s, _bridge)
    S = len(model.scenarios.combinations)
    host_set = {id(r) for r in host_recs}
    components = [c for c in model.flatten_component_tree(rebuild=False) if isinstance(c, Parameter) or id(c) in host_set]
    model.timestepper.reset()
    for t, ts in enumerate(model.timestepper):
        model.timestep = ts
        model.before()
        eng.run(t, 1)
        eng.sync()
        n_failed = eng.program_status()[0]
        if n_failed:
            return
        x, vol, pc = eng.fetch_state()
        for node, cols, w in writers:
            flow = np.zeros(S)
            for c, wc in zip(cols, w):
                flow += wc * x[c]
            _bridge.set_node_flow(node, flow)
        for k, st in enumerate(s.storages):
            _bridge.set_storage_state(st, vol[k], pc[k])
        for c in components:
            c.after()
