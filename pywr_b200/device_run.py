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

from .program import ARRAY_KINDS, Unsupported, _flow_form, compile_program


class DeviceRunResult(dict):
    __getattr__ = dict.__getitem__


def run_on_device(model, engine_factory=None, formulation="route", chunk=None, writeback=True, threads=None,
                  host_mirror=True):
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
    state of a host run; models without them pay nothing."""
    from .bridge import _bridge

    t0 = time.perf_counter()
    if model.dirty or model.timestepper.dirty:
        model.setup()
    else:
        model.reset()
    t1 = time.perf_counter()
    s, prog = compile_program(model, formulation, allow_host_recorders=host_mirror)
    S, T = len(model.scenarios.combinations), prog.n_steps
    host_recs = list(prog.host_recorders)
    if engine_factory is None:
        from .engine import CudaEngine

        def engine_factory(structure, n_scenarios):
            return CudaEngine(structure, n_scenarios, for_program=True)
    eng = engine_factory(s, S)
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
        results = {}
        for r, rec in enumerate(prog.recorders):
            out = eng.fetch_recorder(r, prog.rec_shape(r, S))
            results[rec.name] = out
            if writeback:
                if int(prog.rec_kind[r]) in ARRAY_KINDS:
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
                model.timestepper.reset()
                for ts in model.timestepper:
                    model.timestep = ts
            model.finish()
        t4 = time.perf_counter()
        counters = eng.counters()
        try:
            counters = dict(counters, kernel_variant=eng.stats().get("kernel_variant"))
        except Exception:   # engines without stats() (test doubles)
            pass
    finally:
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
