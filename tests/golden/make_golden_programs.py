"""Generates the device-run golden fixtures (BUILD container only; needs baseline/_ref and
/root/reference).  For each model: (1) a HOST run of the reference framework (Model.run) with the CPU
oracle as solver and the reference's own recorder classes -> expected recorder arrays; (2) the
compiled program + LP structure (pywr_b200.program.compile_program) so that the GPU box can replay
the run through the C-ABI without Pywr.

usage: python tests/golden/make_golden_programs.py
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from pywr.model import Model  # noqa: E402
from pywr.recorders import (DeficitFrequencyNodeRecorder, MeanFlowNodeRecorder, MinimumVolumeStorageRecorder,  # noqa: E402
                            NumpyArrayNodeCurtailmentRatioRecorder, NumpyArrayNodeDeficitRecorder,
                            NumpyArrayNodeRecorder, NumpyArrayNodeSuppliedRatioRecorder, NumpyArrayStorageRecorder,
                            TotalDeficitNodeRecorder, TotalFlowNodeRecorder)

import oracle.pywr_oracle_solver  # noqa: E402,F401
from pywr_b200.program import compile_program  # noqa: E402

REF = os.environ.get("PYWR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    ("thames", "examples/thames/thames.json", "route", "2102-12-31"),
    ("thames", "examples/thames/thames.json", "edge", "2100-12-31"),
    ("two_reservoir", "examples/two_reservoir/two_reservoir.json", "route", "2110-12-31"),
    ("demand_saving2", "tests/models/demand_saving2.json", "route", None),
    ("reservoir2", "tests/models/reservoir2.json", "route", None),
    ("three_storage", "tests/models/three_storage.json", "route", None),
    ("timeseries3", "tests/models/timeseries3.json", "route", None),
    # calendar-driven virtual storages (licences): scheduled resets / activation on the device
    ("seasonal_virtual_storage", "tests/models/seasonal_virtual_storage.json", "route", None),
    ("annual_virtual_storage", "licence:annual", "route", None),
    ("annual_virtual_storage_initial", "licence:annual_initial", "route", None),
    ("monthly_virtual_storage", "licence:monthly", "route", None),
    # rolling licence: per-scenario ring buffer of past use on the device
    ("rolling_virtual_storage", "tests/models/rolling_virtual_storage.json", "route", None),
    ("rolling_virtual_storage_long", "licence:rolling", "route", None),
    # interpolated control curves (ops B200LP_OP_CC_INTERP / CC_PIECEWISE) and an AggregatedStorage on the device
    ("cc_interpolated", "custom:interpolated", "route", "2101-12-31"),
    ("aggregated_storage", "custom:aggregated_storage", "route", None),
    # a reservoir whose COST is an interpolated control curve: the objective changes every timestep (a3, cost_dynamic)
    ("reservoir_with_cc", "tests/models/reservoir_with_cc.json", "route", None),
    # derived parameters: thresholds (B200LP_OP_THRESHOLD) on a parameter, a storage volume, the calendar year and the ordinal
    # day; DivisionParameter, OffsetParameter, ScaledProfileParameter; AggregatedIndexParameter and
    # ConstantScenarioIndexParameter as the index of IndexedArrayParameters; InterpolatedVolumeParameter / InterpolatedParameter
    # (B200LP_OP_INTERP)
    ("derived_parameters", "custom:derived", "route", "2103-12-31"),
    # parameters of YESTERDAY's node flows (op B200LP_OP_LAG on hidden / shared flow recorders): FlowParameter,
    # NodeThresholdParameter, MultipleThresholdIndexParameter
    ("lagged_flows", "custom:lagged", "route", "2102-12-31"),
]


def custom_document(kind):
    import json

    if kind == "interpolated":
        # examples/thames with the demand scaled by a ControlCurveInterpolatedParameter of the reservoir level (values at
        # 100 %, level1, level2, 0 %) and the abstraction capped by a ControlCurvePiecewiseInterpolatedParameter
        doc = json.load(open(os.path.join(REF, "examples/thames/thames.json")))
        for tab in doc.get("tables", {}).values():
            if "url" in tab and not os.path.isabs(tab["url"]):
                tab["url"] = os.path.join(REF, "examples/thames", tab["url"])
        for prm in doc["parameters"].values():
            if isinstance(prm, dict) and "url" in prm and not os.path.isabs(prm["url"]):
                prm["url"] = os.path.join(REF, "examples/thames", prm["url"])
        P = doc["parameters"]
        P["interp_factor"] = {"type": "controlcurveinterpolated", "storage_node": "reservoir1",
                              "control_curves": ["level1", "level2"], "values": [1.0, 0.9, 0.7, 0.4]}
        P["release_cap"] = {"type": "controlcurvepiecewiseinterpolated", "storage_node": "reservoir1",
                            "control_curves": ["level2"], "values": [[3.5, 2.0], [1.5, 0.5]], "minimum": 0.05, "maximum": 0.95}
        P["demand_max_flow"]["parameters"].append("interp_factor")
        next(n for n in doc["nodes"] if n["name"] == "abs1")["max_flow"] = "release_cap"
        return doc
    if kind == "derived":
        import datetime

        doc = json.load(open(os.path.join(REF, "examples/thames/thames.json")))
        for prm in doc["parameters"].values():
            if isinstance(prm, dict) and "url" in prm and not os.path.isabs(prm["url"]):
                prm["url"] = os.path.join(REF, "examples/thames", prm["url"])
        P = doc["parameters"]
        max_volume = next(n for n in doc["nodes"] if n["name"] == "reservoir1")["max_volume"]
        start = datetime.datetime.strptime(str(doc["timestepper"]["start"])[:10], "%Y-%m-%d").date()
        # demand steps up from the second calendar year on, and again whenever the seasonal profile is in its high months
        P["year_step"] = {"type": "currentyearthreshold", "threshold": start.year + 1, "predicate": ">=", "values": [1.0, 1.08]}
        P["summer_extra"] = {"type": "parameterthreshold", "parameter": "demand_profile", "threshold": 1.0, "predicate": "GT",
                             "values": [1.0, 1.05]}
        # restriction when the reservoir holds less than 45 % of its capacity (absolute volume)
        P["low_volume"] = {"type": "storagethreshold", "storage_node": "reservoir1", "threshold": 0.45 * max_volume,
                           "predicate": "<", "values": [1.0, 0.85]}
        P["profile_ratio"] = {"type": "division", "numerator": "level1", "denominator": "level2"}         # 1.4 - 1.86
        P["ratio_offset"] = {"type": "offset", "parameter": "profile_ratio", "offset": -0.6}             # 0.8 - 1.26
        P["scaled"] = {"type": "scaledprofile", "scale": 0.5, "profile": "ratio_offset"}                 # 0.4 - 0.63
        # piecewise-linear curves (B200LP_OP_INTERP): of the stored volume (outside the knots: the fill values) and of a parameter
        P["level_factor"] = {"type": "interpolatedvolume", "node": "reservoir1",
                             "volumes": [0.05 * max_volume, 0.3 * max_volume, 0.7 * max_volume, 0.98 * max_volume],
                             "values": [0.7, 0.9, 0.97, 1.0], "interp_kwargs": {"bounds_error": False, "fill_value": [0.65, 1.01]}}
        P["curve_of_level1"] = {"type": "interpolated", "parameter": "level1", "x": [0.95, 0.6, 0.75], "y": [1.02, 1.0, 0.98],
                                "interp_kwargs": {"bounds_error": False, "fill_value": [1.0, 1.02]}}
        # default bounds_error=True (outside the knots the reference raises, the device run fails the scenario with ST_NAN): on a
        # constant, so that the widened batches of the GPU tests stay inside
        P["curve_of_baseline"] = {"type": "interpolated", "parameter": "demand_baseline", "x": [1.0, 2.0], "y": [0.99, 1.01]}
        # index = first threshold the parameter value reaches (MultipleThresholdParameterIndexParameter)
        P["level1_class"] = {"type": "multiplethresholdparameterindex", "parameter": "level1", "thresholds": [0.9, 0.8, 0.7]}
        P["class_factor"] = {"type": "indexedarray", "index_parameter": "level1_class", "params": [1.0, 0.995, 1.005, 0.99]}
        # the storage's own state as a parameter (B200LP_OP_STATE), polynomials of it and of a parameter, one child per scenario
        # member, a time-only discount factor
        P["stored"] = {"type": "storage", "storage_node": "reservoir1"}
        P["stored_factor"] = {"type": "aggregated", "agg_func": "sum",
                              "parameters": [0.98, {"type": "aggregated", "agg_func": "product", "parameters": [0.0002, "stored"]}]}
        P["poly_of_level"] = {"type": "polynomial1d", "storage_node": "reservoir1", "use_proportional_volume": True,
                              "coefficients": [0.97, 0.08, -0.05], "scale": 0.9, "offset": 0.05}
        P["poly_of_profile"] = {"type": "polynomial1d", "parameter": "demand_profile", "coefficients": [1.0, -0.02, 0.015, 0.001]}
        P["poly_2d"] = {"type": "polynomial2dstorage", "storage_node": "reservoir1", "parameter": "level2", "use_proportional_volume": True,
                        "coefficients": [[1.0, -0.02, 0.01], [0.03, 0.015, -0.01]], "storage_scale": 0.8, "storage_offset": 0.1,
                        "parameter_scale": 1.5, "parameter_offset": -0.2}
        P["by_member"] = {"type": "scenariowrapper", "scenario": "demand",
                          "parameters": [{"type": "constant", "value": 1.0}, "poly_of_profile", {"type": "constant", "value": 0.99},
                                         "curve_of_baseline", {"type": "constant", "value": 1.01}]}
        P["discount"] = {"type": "discountfactor", "rate": 0.035, "base_year": 2100}
        P["demand_max_flow"]["parameters"] += ["year_step", "summer_extra", "low_volume", "level_factor", "curve_of_level1", "curve_of_baseline",
                                               "class_factor", "stored_factor", "poly_of_level", "by_member", "poly_2d"]
        # abstraction capacity: chosen by the sum of two indices (late in the run? storage low?) ...
        P["late"] = {"type": "currentordinaldaythreshold", "threshold": start.toordinal() + 500, "predicate": ">"}
        P["low_volume_index"] = {"type": "storagethreshold", "storage_node": "reservoir1", "threshold": 0.6 * max_volume, "predicate": "LE"}
        P["cap_index"] = {"type": "aggregatedindex", "agg_func": "sum", "parameters": ["late", "low_volume_index"]}
        P["cap_by_state"] = {"type": "indexedarray", "index_parameter": "cap_index", "params": [3.5, 3.0, {"type": "aggregated", "agg_func": "sum", "parameters": [2.0, "scaled"]}]}
        # ... and capped per member of the demand scenario
        P["member_class"] = {"type": "constantscenarioindex", "scenario": "demand", "values": [0, 1, 1, 0, 2]}
        P["cap_by_member"] = {"type": "indexedarray", "index_parameter": "member_class", "params": [3.5, 3.2, 2.4]}
        P["abs_cap"] = {"type": "aggregated", "agg_func": "min", "parameters": [
            "cap_by_state", "cap_by_member", {"type": "aggregated", "agg_func": "product", "parameters": [3.45, "discount"]}]}
        next(n for n in doc["nodes"] if n["name"] == "abs1")["max_flow"] = "abs_cap"
        return doc
    if kind == "lagged":
        doc = json.load(open(os.path.join(REF, "examples/thames/thames.json")))
        for prm in doc["parameters"].values():
            if isinstance(prm, dict) and "url" in prm and not os.path.isabs(prm["url"]):
                prm["url"] = os.path.join(REF, "examples/thames", prm["url"])
        P = doc["parameters"]
        # today's abstraction is capped by yesterday's river flow past the gauge plus an allowance (FlowParameter with an
        # initial value), demand is cut when yesterday's supply was short (NodeThresholdParameter: index 0 on day one) and
        # scaled by a class of yesterday's inflow (MultipleThresholdIndexParameter)
        P["river_yesterday"] = {"type": "flow", "node": "term1", "initial_value": 2.5}
        P["abs_cap"] = {"type": "aggregated", "agg_func": "min",
                        "parameters": [3.5, {"type": "aggregated", "agg_func": "sum", "parameters": ["river_yesterday", 0.8]}]}
        next(n for n in doc["nodes"] if n["name"] == "abs1")["max_flow"] = "abs_cap"
        P["short_yesterday"] = {"type": "nodethreshold", "node": "demand1", "threshold": 1.5, "predicate": "LT", "values": [1.0, 0.9]}
        P["inflow_class"] = {"type": "multiplethresholdindex", "node": "catchment1", "thresholds": [6.0, 3.0, 1.5]}
        P["inflow_factor"] = {"type": "indexedarray", "index_parameter": "inflow_class", "params": [1.05, 1.0, 0.97, 0.92]}
        # yesterday's deficit eases today's demand a little (DeficitParameter), the flow past the gauge three days ago and the
        # mean abstraction of the last week feed the compensation flow (FlowDelayParameter, RollingMeanFlowNodeParameter)
        P["deficit_yesterday"] = {"type": "deficit", "node": "demand1"}
        P["easing"] = {"type": "aggregated", "agg_func": "max",
                       "parameters": [0.9, {"type": "aggregated", "agg_func": "sum", "parameters": [1.0, {"type": "negative", "parameter": "deficit_yesterday"}]}]}
        P["gauge_3_days_ago"] = {"type": "flowdelay", "node": "mrf1", "timesteps": 3, "initial_flow": 0.4}
        P["abs_last_week"] = {"type": "rollingmeanflownode", "node": "abs1", "timesteps": 7, "initial_flow": 1.0}
        P["mrf_today"] = {"type": "aggregated", "agg_func": "sum",
                          "parameters": [0.5, {"type": "aggregated", "agg_func": "product", "parameters": [0.1, "gauge_3_days_ago"]},
                                         {"type": "aggregated", "agg_func": "product", "parameters": [0.05, "abs_last_week"]}]}
        next(n for n in doc["nodes"] if n["name"] == "mrf1")["mrf"] = "mrf_today"
        # a ratchet: once the river past the gauge has run low, a restriction stays for the rest of the run
        P["ever_low"] = {"type": "nodethreshold", "node": "term1", "threshold": 0.9, "predicate": "LT", "values": [1.0, 0.96], "ratchet": True}
        P["demand_max_flow"]["parameters"] += ["short_yesterday", "inflow_factor", "easing", "ever_low"]
        return doc
    if kind == "aggregated_storage":
        # tests/models/three_storage.json: the inflow of reservoir 0 is switched by a control curve on the AGGREGATED storage
        doc = json.load(open(os.path.join(REF, "tests/models/three_storage.json")))
        doc["parameters"] = {"top_up": {"type": "controlcurve", "storage_node": "Total Storage", "control_curves": [0.45, 0.3],
                                        "values": [0.0, 6.0, 14.0]}}
        next(n for n in doc["nodes"] if n["name"] == "Input 0")["max_flow"] = "top_up"
        return doc
    raise KeyError(kind)


def licence_document(kind):
    """Variants of the reference's seasonal_virtual_storage.json with the other calendar-driven licence types."""
    import copy
    import json

    doc = json.load(open(os.path.join(REF, "tests/models/seasonal_virtual_storage.json")))
    lic = doc["nodes"][3]
    if kind.startswith("annual"):
        lic["type"] = "AnnualVirtualStorage"
        lic.pop("end_day"); lic.pop("end_month")
        lic.update(reset_month=4, max_volume=1500, initial_volume=900)
        doc["timestepper"]["end"] = "2017-06-30"
        if kind == "annual_initial":
            lic["reset_to_initial_volume"] = True
    elif kind == "rolling":   # 30-day rolling licence smaller than the demand over the window
        lic["type"] = "RollingVirtualStorage"
        for k in ("end_day", "end_month", "reset_day", "reset_month"):
            lic.pop(k)
        lic.update(days=30, max_volume=200, initial_volume=120)
    elif kind == "monthly":
        lic["type"] = "MonthlyVirtualStorage"
        for k in ("end_day", "end_month", "reset_day", "reset_month"):
            lic.pop(k)
        lic.update(months=2, initial_months=1, max_volume=400, initial_volume=250)
    return copy.deepcopy(doc)


def add_recorders(m):
    recs = []
    for n in list(m.nodes):
        cn = type(n).__name__
        if cn in ("Storage", "Reservoir", "AnnualVirtualStorage", "SeasonalVirtualStorage", "MonthlyVirtualStorage",
                  "RollingVirtualStorage") or (cn == "AggregatedStorage" and "aggregated_storage" in str(m.metadata.get("b200_case", ""))):
            recs += [NumpyArrayStorageRecorder(m, n, name=f"vol_{n.name}"),
                     NumpyArrayStorageRecorder(m, n, proportional=True, name=f"pc_{n.name}"),
                     MinimumVolumeStorageRecorder(m, n, name=f"minv_{n.name}")]
        elif cn in ("Output", "Link", "Input", "Catchment", "River", "RiverGauge"):
            recs.append(NumpyArrayNodeRecorder(m, n, name=f"flow_{n.name}"))
            if cn == "Output":
                recs += [NumpyArrayNodeDeficitRecorder(m, n, name=f"def_{n.name}"),
                         TotalDeficitNodeRecorder(m, n, name=f"tdef_{n.name}"),
                         NumpyArrayNodeSuppliedRatioRecorder(m, n, name=f"sr_{n.name}"),
                         NumpyArrayNodeCurtailmentRatioRecorder(m, n, name=f"cr_{n.name}"),
                         TotalFlowNodeRecorder(m, n, name=f"tf_{n.name}", factor=2.0),
                         MeanFlowNodeRecorder(m, n, name=f"mf_{n.name}"),
                         DeficitFrequencyNodeRecorder(m, n, name=f"df_{n.name}")]
    return recs


def load(path, formulation, end):
    if path.startswith("custom:"):
        src = custom_document(path.split(":")[1])
    else:
        src = licence_document(path.split(":")[1]) if path.startswith("licence:") else os.path.join(REF, path)
    m = Model.load(src, solver="oracle" if formulation == "route" else "oracle-edge")
    if end:
        m.timestepper.end = end
    for r in list(m.recorders):  # keep only the recorders added below
        pass
    return m


def run_case(name, path, formulation, end):
    m = load(path, formulation, end)
    m.metadata["b200_case"] = name
    recs = add_recorders(m)
    m.run()
    expected = {r.name: (np.array(r.data) if hasattr(r, "data") else np.array(r.values())) for r in recs}
    m2 = load(path, formulation, end)
    m2.metadata["b200_case"] = name
    recs2 = add_recorders(m2)
    m2.setup()
    s, prog = compile_program(m2, formulation, recorders=recs2)
    base = os.path.join(OUT, f"{name}.{formulation}")
    s.save(base + ".pstructure.npz")
    prog.save(base + ".program.npz")
    np.savez_compressed(base + ".expected.npz", names=np.array([r.name for r in prog.recorders]),
                        **{f"rec{i}": expected[r.name] for i, r in enumerate(prog.recorders)})
    print(f"{name}.{formulation}: T={prog.n_steps} S={len(m2.scenarios.combinations)} ops={prog.n_ops} "
          f"tables={prog.n_tables} recorders={prog.n_recorders} slots={s.n_slots}")


if __name__ == "__main__":
    only = sys.argv[1:]
    for case in CASES:
        if not only or case[0] in only:
            run_case(*case)
