"""Benchmark workloads (BASELINE.json `configs`): model documents, the
synthetic stochastic inflow generator, and run-time construction of device programs.

Two ways to obtain the program of a workload:
  * ``build_model(...)`` + ``pywr_b200.program.compile_program`` — needs the reference framework
    (baseline/_ref); this is what a user does and what ``benchmarks/make_fixtures.py`` runs;
  * ``load_workload(...)`` — loads the structure + single-scenario program compiled by
    make_fixtures.py (committed under benchmarks/fixtures/) and widens it to S scenarios by swapping in
    the synthetic inflow table; needs only numpy, so ``bench.py`` runs on a box without Pywr.
``tests/test_workloads.py`` checks that both give the same program.
"""

from __future__ import annotations

import copy
import os

import numpy as np

FIXTURES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures")

# Summary statistics of column `flow` of the reference's examples/data/thames_stochastic_flow.gz
# (36,524 daily values, m3/s): mean of log(flow) per calendar month, standard deviation and lag-1
# autocorrelation of the de-seasonalised log flow.  Derived numbers, not the data.
THAMES_LOG_MEAN_BY_MONTH = np.array([3.8735, 4.2096, 4.2449, 3.8334, 3.3088, 2.9321, 2.7325, 2.7017, 2.7557,
                                     2.9138, 3.0789, 3.3541])
THAMES_LOG_RESID_STD = 0.7006
FLOW_FACTOR = 0.0864  # m3/s -> Ml/d, the models' `flow_factor` parameter

SEEDS = {"two_reservoir": 20240901, "thames": 20240902, "two_reservoir_params": 20240904}


def daily_index(start, years=None, end=None):
    import pandas as pd

    if end is None:
        end = pd.Timestamp(start) + pd.DateOffset(years=years) - pd.Timedelta(days=1)
    return pd.date_range(start, end, freq="D")


def synthetic_inflow(months, n_scenarios, seed, first_scenario=0, rho=0.9):
    """[T, n_scenarios] float64 daily inflow in model units.  Log-space AR(1) with lag-1 correlation
    `rho` around the monthly log-mean of the Thames record, independent per scenario; scenario k of the
    GLOBAL batch always gets the same series whatever the number of ranks (its generator is seeded
    with (seed, k))."""
    months = np.asarray(months)
    T = months.shape[0]
    mu = THAMES_LOG_MEAN_BY_MONTH[months - 1]
    out = np.empty((T, n_scenarios))
    sd = THAMES_LOG_RESID_STD
    innov = sd * np.sqrt(1.0 - rho * rho)
    try:
        from scipy.signal import lfilter
    except Exception:  # pragma: no cover
        lfilter = None
    for k in range(n_scenarios):
        rng = np.random.default_rng([seed, first_scenario + k])
        e = rng.standard_normal(T) * innov
        e[0] = rng.standard_normal() * sd
        if lfilter is not None:
            z = lfilter([1.0], [1.0, -rho], e)
        else:
            z = np.empty(T)
            z[0] = e[0]
            for t in range(1, T):
                z[t] = rho * z[t - 1] + e[t]
        out[:, k] = np.exp(mu + z) * FLOW_FACTOR
    return out


# ---- model documents (same topology and numbers as the reference's example models, restated) -----
def two_reservoir_document(start, end, control_curve=0.5):
    """examples/two_reservoir: two catchments feeding two reservoirs with compensation releases
    (river gauges), two demands and a controlled transfer from reservoir2 to reservoir1."""
    return {
        "metadata": {"title": "two reservoir (benchmark)", "minimum_version": "0.1"},
        "timestepper": {"start": start, "end": end, "timestep": 1},
        "nodes": [
            {"name": "catchment1", "type": "catchment", "flow": 0.0},
            {"name": "catchment2", "type": "catchment", "flow": 0.0},
            {"name": "reservoir1", "type": "storage", "min_volume": 0, "max_volume": 200, "initial_volume": 200, "cost": -5},
            {"name": "reservoir2", "type": "storage", "min_volume": 0, "max_volume": 200, "initial_volume": 200, "cost": -5},
            {"name": "transfer", "type": "link", "cost": -500, "max_flow": "transfer_controller"},
            {"name": "demand1", "type": "output", "max_flow": 1.8, "cost": -101},
            {"name": "demand2", "type": "output", "max_flow": 1.0, "cost": -100},
            {"name": "river1", "type": "link"},
            {"name": "river2", "type": "link"},
            {"name": "compensation1", "type": "rivergauge", "mrf": 0.6, "mrf_cost": -1000},
            {"name": "compensation2", "type": "rivergauge", "mrf": 0.6, "mrf_cost": -1000},
            {"name": "terminator", "type": "output"},
        ],
        "edges": [
            ["catchment1", "reservoir1"], ["catchment2", "reservoir2"], ["reservoir1", "river1"],
            ["reservoir2", "river2"], ["river1", "compensation1"], ["river2", "compensation2"],
            ["compensation1", "terminator"], ["compensation2", "terminator"], ["reservoir2", "transfer"],
            ["transfer", "reservoir1"], ["reservoir1", "demand1"], ["reservoir2", "demand2"],
        ],
        "parameters": {
            "control_curve": {"type": "monthlyprofile", "values": [control_curve] * 12},
            "transfer_controller": {"type": "controlcurve", "storage_node": "reservoir1",
                                    "control_curve": "control_curve", "values": [0.0, 1.0]},
        },
        "recorders": {
            "deficit1": {"type": "DeficitFrequencyNode", "node": "demand1"},
            "deficit2": {"type": "DeficitFrequencyNode", "node": "demand2"},
            "transferred": {"type": "totalflownode", "node": "transfer"},
            "volume1": {"type": "NumpyArrayStorageRecorder", "node": "reservoir1"},
            "volume2": {"type": "NumpyArrayStorageRecorder", "node": "reservoir2"},
            "supplied1": {"type": "NumpyArrayNodeRecorder", "node": "demand1"},
            "supplied2": {"type": "NumpyArrayNodeRecorder", "node": "demand2"},
        },
    }


def thames_document(start, end):
    """examples/thames: catchment -> river -> (minimum residual flow gauge -> sea | abstraction ->
    reservoir -> demand) with demand saving driven by the reservoir's control curves; the shipped
    5-member `demand` scenario (growth factors) is kept and sliced to its central member
    (`slice: [2, 3]`, pywr/_core.pyx:122-124)."""
    prof = lambda a, b: [a, a, a, a, b, b, b, b, a, a, a, a]  # noqa: E731
    return {
        "metadata": {"title": "thames (benchmark)", "minimum_version": "0.1"},
        "timestepper": {"start": start, "end": end, "timestep": 1},
        "scenarios": [{"name": "demand", "size": 5, "slice": [2, 3]}],
        "nodes": [
            {"name": "catchment1", "type": "catchment", "flow": 0.0},
            {"name": "river1", "type": "link"},
            {"name": "mrf1", "type": "rivergauge", "mrf": 0.6, "mrf_cost": -1000},
            {"name": "term1", "type": "output"},
            {"name": "reservoir1", "type": "storage", "max_volume": 200, "initial_volume": 200, "cost": -10},
            {"name": "abs1", "type": "link", "max_flow": 3.5},
            {"name": "demand1", "type": "output", "max_flow": "demand_max_flow", "cost": -500},
        ],
        "edges": [["catchment1", "river1"], ["river1", "mrf1"], ["mrf1", "term1"], ["river1", "abs1"],
                  ["abs1", "reservoir1"], ["reservoir1", "demand1"]],
        "parameters": {
            "demand_baseline": {"type": "constant", "value": 1.7},
            "demand_profile": {"type": "monthlyprofile", "values": prof(0.9, 1.2)},
            "demand_max_flow": {"type": "aggregated", "agg_func": "product",
                                "parameters": ["demand_baseline", "demand_profile", "demand_saving_factor",
                                               "demand_growth_factor"]},
            "level1": {"type": "monthlyprofile", "values": [0.8, 0.85, 0.9, 0.92, 0.9, 0.85, 0.8, 0.7, 0.65, 0.7, 0.75, 0.8]},
            "level2": {"type": "monthlyprofile", "values": [0.5, 0.55, 0.6, 0.62, 0.6, 0.55, 0.5, 0.4, 0.35, 0.4, 0.45, 0.5]},
            "demand_saving_level": {"type": "controlcurveindex", "storage_node": "reservoir1",
                                    "control_curves": ["level1", "level2"]},
            "demand_saving_factor": {"type": "indexedarray", "index_parameter": "demand_saving_level",
                                     "params": ["level0_factor", "level1_factor", "level2_factor"]},
            "level0_factor": {"type": "constant", "value": 1.0},
            "level1_factor": {"type": "monthlyprofile", "values": prof(0.95, 0.9)},
            "level2_factor": {"type": "monthlyprofile", "values": prof(0.8, 0.75)},
            "demand_growth_factor": {"type": "constantscenario", "scenario": "demand",
                                     "values": [0.75, 0.9, 1.0, 1.1, 1.25]},
        },
        "recorders": {
            "volume": {"type": "NumpyArrayStorageRecorder", "node": "reservoir1"},
            "supplied": {"type": "NumpyArrayNodeRecorder", "node": "demand1"},
            "deficit": {"type": "TotalDeficitNodeRecorder", "node": "demand1"},
        },
    }


WORKLOADS = {
    # name: (document, start, years, inflow catchments, seed key)
    "two_reservoir": dict(document=two_reservoir_document, start="2100-01-01", years=50,
                          catchments=("catchment1", "catchment2"), seed="two_reservoir"),
    "thames": dict(document=thames_document, start="2100-01-01", years=30, catchments=("catchment1",), seed="thames"),
}


def build_model(name, n_scenarios, years=None, solver=None, inflow=None, first_scenario=0, solver_args=None):
    """The workload as a live Pywr model (needs the reference framework): the document above plus a
    Scenario("inflow", size=S) and an ArrayIndexedScenarioParameter with the synthetic inflow feeding
    every catchment."""
    from pywr.core import Scenario
    from pywr.model import Model
    from pywr.parameters import ArrayIndexedScenarioParameter

    w = WORKLOADS[name]
    years = years or w["years"]
    idx = daily_index(w["start"], years=years)
    doc = w["document"](str(idx[0].date()), str(idx[-1].date()))
    kwargs = {"solver": solver} if solver else {}
    if solver_args:
        kwargs["solver_args"] = dict(solver_args)
    model = Model.load(doc, **kwargs)
    scenario = Scenario(model, "inflow", size=n_scenarios)
    if inflow is None:
        inflow = synthetic_inflow(idx.month.values, n_scenarios, SEEDS[w["seed"]], first_scenario)
    param = ArrayIndexedScenarioParameter(model, scenario, inflow, name="inflow")
    for c in w["catchments"]:
        model.nodes[c].flow = param
    return model


def build_model_param_sets(n_param_sets, n_inflow, years=30, solver=None, seed=None):
    """BASELINE.json config 5 / SURVEY.md 8(d): the optimisation-style batch.  two_reservoir with two
    scenario axes, "params" (one control-curve monthly profile per candidate parameter set, the outer
    loop of the reference's optimisation wrappers, pywr/optimisation/platypus.py:115-154, flattened
    into scenarios) x "inflow"; S = n_param_sets * n_inflow combinations in Pywr's own order
    (pywr/_core.pyx:210-240: the last axis varies fastest)."""
    from pywr.core import Scenario
    from pywr.model import Model
    from pywr.parameters import ArrayIndexedScenarioParameter, ScenarioMonthlyProfileParameter

    w = WORKLOADS["two_reservoir"]
    idx = daily_index(w["start"], years=years)
    doc = w["document"](str(idx[0].date()), str(idx[-1].date()))
    del doc["parameters"]["control_curve"]
    doc["parameters"]["transfer_controller"]["control_curve"] = 0.5   # replaced below
    kwargs = {"solver": solver} if solver else {}
    model = Model.load(doc, **kwargs)
    params = Scenario(model, "params", size=n_param_sets)
    inflow_axis = Scenario(model, "inflow", size=n_inflow)
    rng = np.random.default_rng(SEEDS["two_reservoir_params"] if seed is None else seed)
    curve = ScenarioMonthlyProfileParameter(model, params, rng.uniform(0.0, 1.0, size=(n_param_sets, 12)), name="control_curve")
    model.parameters["transfer_controller"].control_curves = [curve]
    inflow = synthetic_inflow(idx.month.values, n_inflow, SEEDS["two_reservoir"])
    flow = ArrayIndexedScenarioParameter(model, inflow_axis, inflow, name="inflow")
    for c in w["catchments"]:
        model.nodes[c].flow = flow
    return model


def synthetic_network_document(n_reaches=12, seed=20240903, start="2100-01-01", end="2100-03-31"):
    """A generated river network (scaled by `n_reaches`): a main stem of river
    reaches, each with a catchment inflow, some with a tributary reservoir (storage + demand + spill
    back to the stem), abstraction demands, a minimum-residual-flow gauge every few reaches, one
    aggregated node capping total abstraction and a transfer between two reservoirs.  Meant for the
    edge formulation (routes multiply with every confluence)."""
    rng = np.random.default_rng(seed)
    nodes, edges, params = [], [], {}
    prev = None
    reservoirs, abstractions = [], []
    for r in range(n_reaches):
        c, link = f"catch{r:03d}", f"reach{r:03d}"
        params[f"inflow{r:03d}"] = {"type": "monthlyprofile", "values": [float(v) for v in rng.uniform(2.0, 12.0, 12)]}
        nodes.append({"name": c, "type": "catchment", "flow": f"inflow{r:03d}"})
        nodes.append({"name": link, "type": "link"})
        edges.append([c, link])
        if prev is not None:
            edges.append([prev, link])
        prev = link
        if r % 3 == 1:  # tributary reservoir fed from the reach, supplying a demand, spilling downstream
            res, dem, ab = f"res{r:03d}", f"town{r:03d}", f"abs{r:03d}"
            nodes.append({"name": ab, "type": "link", "max_flow": float(rng.uniform(3.0, 8.0)), "cost": 1.0})
            nodes.append({"name": res, "type": "storage", "max_volume": float(rng.uniform(80, 300)),
                          "initial_volume_pc": 0.8, "cost": -float(rng.uniform(2, 20))})
            nodes.append({"name": dem, "type": "output", "max_flow": float(rng.uniform(2.0, 7.0)), "cost": -float(rng.uniform(200, 600))})
            edges += [[link, ab], [ab, res], [res, dem]]
            reservoirs.append(res); abstractions.append(ab)
        if r % 4 == 2:  # direct river abstraction
            dem = f"farm{r:03d}"
            nodes.append({"name": dem, "type": "output", "max_flow": float(rng.uniform(0.5, 3.0)), "cost": -float(rng.uniform(50, 150))})
            edges.append([link, dem])
        if r % 5 == 4:  # gauge with a minimum residual flow on the stem
            g = f"gauge{r:03d}"
            nodes.append({"name": g, "type": "rivergauge", "mrf": float(rng.uniform(1.0, 4.0)), "mrf_cost": -1000.0})
            edges.append([link, g])
            prev = g
    nodes.append({"name": "sea", "type": "output"})
    edges.append([prev, "sea"])
    if len(reservoirs) >= 2:
        nodes.append({"name": "transfer", "type": "link", "max_flow": 2.0, "cost": 5.0})
        edges += [[reservoirs[-1], "transfer"], ["transfer", reservoirs[0]]]
    if len(abstractions) >= 2:
        nodes.append({"name": "licence", "type": "aggregatednode", "nodes": abstractions, "max_flow": float(4.0 * len(abstractions))})
    return {"metadata": {"title": "generated network", "minimum_version": "0.1"},
            "timestepper": {"start": start, "end": end, "timestep": 1},
            "nodes": nodes, "edges": edges, "parameters": params, "recorders": {}}


def load_workload(name, n_scenarios, years=None, first_scenario=0, inflow=None):
    """(structure, program, info) for S scenarios from the committed single-scenario fixture."""
    from pywr_b200.program import Program
    from pywr_b200.structure import LPStructure

    w = WORKLOADS[name]
    s = LPStructure.load(os.path.join(FIXTURES, f"{name}.pstructure.npz"))
    base = Program.load(os.path.join(FIXTURES, f"{name}.program.npz"))
    years = years or w["years"]
    idx = daily_index(w["start"], years=years)
    T = len(idx)
    if T > base.n_steps:
        raise ValueError(f"fixture holds {base.n_steps} timesteps, {T} requested")
    S = int(n_scenarios)
    if inflow is None:
        inflow = synthetic_inflow(idx.month.values, S, SEEDS[w["seed"]], first_scenario)
    assert inflow.shape == (T, S)
    q = copy.copy(base)
    q.n_steps = T
    q.ts_days = base.ts_days[:T].copy()
    rows, cols, axis, offs, chunks, pos = [], [], [], [], [], 0
    for t in range(base.n_tables):
        r0, c0 = int(base.table_rows[t]), int(base.table_cols[t])
        data = base.table_data[base.table_offset[t]: base.table_offset[t] + r0 * c0].reshape(r0, c0)
        if base.table_axis[t] >= 0 and r0 > 1:      # the inflow table ([T, members] over scenario axis "inflow")
            data = inflow
        elif r0 > 1:
            data = data[:T]
        rows.append(data.shape[0]); cols.append(data.shape[1]); axis.append(int(base.table_axis[t]))
        offs.append(pos); chunks.append(np.ascontiguousarray(data).reshape(-1)); pos += data.size
    q.table_rows, q.table_cols = np.array(rows, np.int32), np.array(cols, np.int32)
    q.table_axis, q.table_offset = np.array(axis, np.int32), np.array(offs, np.int64)
    q.table_data = np.concatenate(chunks)
    # scenario indices: the inflow axis (the one the [T, S] table is indexed by) counts the scenarios of this
    # batch, every other axis keeps the member the fixture was compiled with (thames: `demand` sliced to member 2)
    n_ax = max(base.n_axes, 1)
    inflow_axes = {int(a) for a, r in zip(base.table_axis, base.table_rows) if a >= 0 and r > 1}
    base_idx = np.asarray(base.scen_index, np.int32).reshape(n_ax, -1)[:, 0]
    q.scen_index = np.stack([np.arange(S, dtype=np.int32) if (a in inflow_axes or base.n_axes == 0) else
                             np.full(S, base_idx[a], np.int32) for a in range(n_ax)]).reshape(-1)
    nst = max(s.n_storage, 1)
    q.initial_volume = np.repeat(base.initial_volume.reshape(nst, -1)[:, :1], S, axis=1).reshape(-1)
    q.initial_pc = np.repeat(base.initial_pc.reshape(nst, -1)[:, :1], S, axis=1).reshape(-1)
    info = dict(name=name, n_scenarios=S, timesteps=T, rows=s.n_rows, cols=s.n_cols, nnz=s.nnz, nodes=s.n_nodes,
                n_storage=s.n_storage, n_recorders=q.n_recorders,
                n_array_recorders=int(np.sum((q.rec_kind >= 1) & (q.rec_kind <= 6))))
    return s, q, info


def algorithmic_bytes_per_scenario_step(info):
    """B = 8(n + 2m) + 2(m + n) + 16 n_st + 8 N + 8 n_rec."""
    m, n = info["rows"], info["cols"]
    return 8 * (n + 2 * m) + 2 * (m + n) + 16 * info["n_storage"] + 8 * info["nodes"] + 8 * info["n_recorders"]


# ---- SURVEY.md 8(d) config 4 as specified ----------------------------------------------------------------------------
def spec_network_document(scale=1.0, seed=20240903, start="2100-01-01", end="2100-12-31"):
    """The generated network of config 4 with SURVEY 8(d)'s mix (at scale 1.0: ~2,000 primitive nodes — 150 reservoirs, 250
    catchments, 300 demands, links for the rest — 40 AggregatedNodes, half of them with factors, 60 PiecewiseLinks of 2-3 steps,
    river-TREE topology plus 4 transfers that close cycles; edge formulation).  Catchment inflows are monthly profiles here;
    ``add_stochastic_inflows`` multiplies them by per-scenario daily AR(1) series.

    Topology: reach k (k > 0) flows into an earlier reach chosen near k, so the tree is a few dozen reaches deep; leaves and
    some inner reaches carry a catchment; every few reaches a tributary reservoir (reach -> abstraction -> storage -> town),
    a direct river abstraction, or a minimum-residual-flow gauge; 60 reaches are PiecewiseLinks (a cheap channel of limited
    capacity and a costly overflow); licences (AggregatedNode max_flow over ~6 abstractions) and paired intakes whose flows must
    keep a fixed ratio (AggregatedNode factors); transfers take water from a reservoir back to a reach upstream of its own
    intake (the only cycles)."""
    rng = np.random.default_rng(seed)
    n_reach = max(12, int(round(430 * scale)))
    n_res, n_catch, n_farm = max(3, int(round(150 * scale))), max(6, int(round(250 * scale))), max(3, int(round(150 * scale)))
    n_gauge, n_pw = max(2, int(round(35 * scale))), max(2, int(round(60 * scale)))
    n_lic, n_fac, n_transfer = max(1, int(round(20 * scale))), max(1, int(round(20 * scale))), min(4, max(1, int(round(4 * scale))))
    nodes, edges, params = [], [], {}
    # river tree: reach k drains into parent[k] < k; the outlet is reach 0
    parent = [-1] + [int(max(0, k - 1 - rng.integers(0, min(k, 12)))) for k in range(1, n_reach)]
    children = [[] for _ in range(n_reach)]
    for k in range(1, n_reach):
        children[parent[k]].append(k)
    leaves = [k for k in range(n_reach) if not children[k]]
    inner = [k for k in range(1, n_reach) if children[k]]
    catch_at = list(leaves)[:n_catch]
    if len(catch_at) < n_catch:
        catch_at += [int(k) for k in rng.choice(inner, size=min(len(inner), n_catch - len(catch_at)), replace=False)]
    pw_at = set(int(k) for k in rng.choice(np.arange(1, n_reach), size=min(n_pw, n_reach - 1), replace=False))
    gauge_at = set(int(k) for k in rng.choice([k for k in range(1, n_reach) if k not in pw_at], size=n_gauge, replace=False))
    reach_name = [f"reach{k:04d}" for k in range(n_reach)]
    for k in range(n_reach):
        if k in pw_at:   # cheap channel of limited capacity, then a costly overflow (2 or 3 steps)
            steps = int(rng.integers(2, 4))
            caps = sorted(float(v) for v in rng.uniform(20.0, 120.0, steps - 1))
            nodes.append({"name": reach_name[k], "type": "piecewiselink", "nsteps": steps, "max_flows": caps + [None],
                          "costs": [0.0] + [float(v) for v in np.sort(rng.uniform(0.5, 3.0, steps - 1))]})
        else:
            nodes.append({"name": reach_name[k], "type": "link"})
    out_of = {}   # what flows on downstream from reach k (the reach itself, or its gauge)
    for k in range(n_reach):
        out_of[k] = reach_name[k]
        if k in gauge_at:
            g = f"gauge{k:04d}"
            nodes.append({"name": g, "type": "rivergauge", "mrf": float(rng.uniform(0.5, 3.0)), "mrf_cost": -1000.0})
            edges.append([reach_name[k], g])
            out_of[k] = g
    for k in range(1, n_reach):
        edges.append([out_of[k], reach_name[parent[k]]])
    nodes.append({"name": "sea", "type": "output"})
    edges.append([out_of[0], "sea"])
    for i, k in enumerate(catch_at):
        c = f"catch{i:04d}"
        params[f"inflow{i:04d}"] = {"type": "monthlyprofile", "values": [float(v) for v in rng.uniform(1.0, 9.0, 12)]}
        nodes.append({"name": c, "type": "catchment", "flow": f"inflow{i:04d}"})
        edges.append([c, reach_name[k]])
    # tributary reservoirs and direct abstractions on distinct reaches
    sites = [int(k) for k in rng.permutation(np.arange(1, n_reach))]
    res_at, farm_at = sites[:n_res], sites[n_res:n_res + n_farm]
    abstractions, reservoirs = [], []
    for i, k in enumerate(res_at):
        ab, res, town = f"abs{i:04d}", f"res{i:04d}", f"town{i:04d}"
        nodes.append({"name": ab, "type": "link", "max_flow": float(rng.uniform(3.0, 9.0)), "cost": 1.0})
        nodes.append({"name": res, "type": "storage", "max_volume": float(rng.uniform(80, 400)), "initial_volume_pc": 0.8,
                      "cost": -float(rng.uniform(2, 20))})
        nodes.append({"name": town, "type": "output", "max_flow": float(rng.uniform(2.0, 7.0)), "cost": -float(rng.uniform(200, 600))})
        edges += [[reach_name[k], ab], [ab, res], [res, town]]
        abstractions.append(ab); reservoirs.append((res, k))
    farms = []
    for i, k in enumerate(farm_at):
        intake, farm = f"intake{i:04d}", f"farm{i:04d}"
        nodes.append({"name": intake, "type": "link", "max_flow": float(rng.uniform(0.5, 4.0))})
        nodes.append({"name": farm, "type": "output", "max_flow": float(rng.uniform(0.5, 3.0)), "cost": -float(rng.uniform(50, 150))})
        edges += [[reach_name[k], intake], [intake, farm]]
        farms.append(intake)
    # transfers: reservoir -> a reach upstream of its own intake (cycle), with a cost so that nothing circulates for free
    upstream_of = lambda k: [c for c in children[k]] or [k]  # noqa: E731
    for i in range(n_transfer):
        res, k = reservoirs[int(rng.integers(0, len(reservoirs)))]
        dest = upstream_of(k)[0]
        t = f"transfer{i}"
        nodes.append({"name": t, "type": "link", "max_flow": 2.0, "cost": 5.0})
        if i % 2 == 0:   # from the reservoir (a cycle through storage: water can go round over several timesteps)
            edges += [[res, t], [t, reach_name[dest]]]
        else:            # pumped straight from the reach back upstream: a cycle of the graph itself
            edges += [[reach_name[k], t], [t, reach_name[dest]]]
    # licences: total abstraction of a group
    pool = list(abstractions)
    rng.shuffle(pool)
    for i in range(n_lic):
        group = pool[i * 6:(i + 1) * 6] or pool[:3]
        nodes.append({"name": f"licence{i:02d}", "type": "aggregatednode", "nodes": group, "max_flow": float(3.5 * len(group))})
    # paired intakes that must keep a fixed ratio
    fp = list(farms)
    rng.shuffle(fp)
    for i in range(min(n_fac, len(fp) // 2)):
        a, b = fp[2 * i], fp[2 * i + 1]
        nodes.append({"name": f"ratio{i:02d}", "type": "aggregatednode", "nodes": [a, b], "factors": [1.0, float(rng.choice([0.5, 1.0, 2.0]))]})
    return {"metadata": {"title": "generated network (SURVEY 8d config 4 mix)", "minimum_version": "0.1"},
            "timestepper": {"start": start, "end": end, "timestep": 1},
            "nodes": nodes, "edges": edges, "parameters": params, "recorders": {}}


STOCHASTIC_SEED = 20240905


def regional_wetness(T, n_scenarios, region, seed=STOCHASTIC_SEED, first_scenario=0, rho=0.9, sigma=0.45):
    """[T, n_scenarios] log-normal AR(1) wetness factor of one region (mean 1): scenario k of the GLOBAL batch always gets the
    same series (its generator is seeded with (seed, region, k)), whatever the batch size or the number of ranks."""
    innov = sigma * np.sqrt(1.0 - rho * rho)
    e = np.empty((T, n_scenarios))
    z0 = np.empty(n_scenarios)
    for k in range(n_scenarios):
        rng = np.random.default_rng([seed, region, first_scenario + k])
        e[:, k] = rng.standard_normal(T) * innov
        z0[k] = rng.standard_normal() * sigma
    z = np.empty((T, n_scenarios))
    z[0] = z0
    for t in range(1, T):
        z[t] = rho * z[t - 1] + e[t]
    return np.exp(z - 0.5 * sigma * sigma)


def add_stochastic_inflows(model, n_scenarios, n_regions=4, seed=STOCHASTIC_SEED, first_scenario=0, rho=0.9, sigma=0.45):
    """Scenario axis "inflow" with per-scenario DAILY stochastic inflows: every catchment's monthly profile is
    multiplied by the log-normal AR(1) wetness series of its region (n_regions series per scenario, [T, S] each; scenario k of
    the global batch always gets the same series).  All of it is device-evaluable (tables indexed by the scenario axis +
    product ops)."""
    from pywr.core import Scenario
    from pywr.parameters import AggregatedParameter, ArrayIndexedScenarioParameter

    T = len(model.timestepper) if model.timestepper._periods is not None else None
    if T is None:
        model.timestepper.setup()
        T = len(model.timestepper)
    axis = Scenario(model, "inflow", size=n_scenarios)
    regions = [ArrayIndexedScenarioParameter(model, axis, regional_wetness(T, n_scenarios, r, seed, first_scenario, rho, sigma),
                                             name=f"wetness{r}") for r in range(n_regions)]
    catchments = sorted((n for n in model.nodes if n.name.startswith("catch")), key=lambda n: n.name)
    for i, node in enumerate(catchments):
        node.flow = AggregatedParameter(model, [node.max_flow, regions[i % n_regions]], agg_func="product", name=f"stoch_{node.name}")
    return axis
