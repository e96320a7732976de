# cython: language_level=3, boundscheck=False, wraparound=False
"""Cython bridge between the device-resident run and Pywr's own objects.

Pywr keeps the per-scenario state the solver path touches in NON-public ``cdef`` members
(``AbstractNode._flow/_prev_flow`` pywr/_core.pxd:47-48, ``NumpyArrayNodeRecorder._data``
pywr/recorders/_recorders.pxd:43, ``BaseConstantNodeRecorder._values`` :116, the value arrays of the
array-backed Parameters pywr/parameters/_parameters.pxd:36-84).  They are reachable only from a
Cython module that ``cimport``s Pywr's pxd files (installed with the package, setup.py:124-131) and
is compiled against the same build — this module.  It (a) reads the data of array-backed Parameters
so that they can be uploaded once, and (b) writes device results back into the reference's own
recorder / node objects so that ``recorder.data``, ``values()``, ``aggregated_value()``,
``to_dataframe()`` and ``node.flow`` behave exactly as after a host run.
"""
import numpy as np
cimport numpy as np

from pywr._core cimport AbstractNode, AbstractStorage, RollingVirtualStorage
from pywr.parameters._parameters cimport (
    Parameter, ConstantScenarioParameter, ArrayIndexedParameter, ArrayIndexedScenarioParameter,
    DataFrameParameter, ScenarioMonthlyProfileParameter, ScenarioDailyProfileParameter,
    ScenarioWeeklyProfileParameter, ArrayIndexedScenarioMonthlyFactorsParameter, ConstantScenarioIndexParameter,
)
from pywr.parameters._thresholds cimport (AbstractThresholdParameter, MultipleThresholdParameterIndexParameter,
                                          MultipleThresholdIndexParameter)
from pywr.parameters._polynomial cimport Polynomial1DParameter, Polynomial2DStorageParameter
from pywr.recorders._recorders cimport (
    NumpyArrayNodeRecorder, NumpyArrayAbstractStorageRecorder, BaseConstantNodeRecorder, NumpyArrayParameterRecorder,
    BaseConstantParameterRecorder,
    NumpyArrayIndexParameterRecorder,
    BaseConstantStorageRecorder, NodeRecorder, StorageRecorder,
)


def parameter_arrays(Parameter p):
    """Data of the array-backed, state-independent Parameter classes.

    Returns ``None`` for other classes, else a dict with ``values`` ([rows, cols] float64, rows = 1
    or the number of timesteps), ``axis`` (scenario axis the columns follow, or None), ``offset``
    (DataFrameParameter.timestep_offset) and ``col_map`` (int32 map from scenario member to column
    when the data was sliced, else None).
    """
    cdef DataFrameParameter df
    cdef ArrayIndexedParameter ai
    cdef ArrayIndexedScenarioParameter ais
    cdef ConstantScenarioParameter cs
    if type(p).__name__ == "DataFrameParameter":
        df = p
        if df.scenario is None:
            return dict(values=np.asarray(df._values)[:, :1].copy(), axis=None, offset=df.timestep_offset, col_map=None)
        # a view, not a copy: the caller packs it into the program's table block (one copy of a [T, S] array instead of two)
        return dict(values=np.asarray(df._values), axis=df._scenario_index, offset=df.timestep_offset,
                    col_map=None if df._scenario_ids is None else np.asarray(df._scenario_ids).copy())
    if type(p).__name__ == "ArrayIndexedParameter":
        ai = p
        return dict(values=np.asarray(ai.values).reshape(-1, 1).copy(), axis=None, offset=0, col_map=None)
    if type(p).__name__ == "ArrayIndexedScenarioParameter":
        ais = p
        return dict(values=np.asarray(ais.values), axis=ais._scenario_index, offset=0, col_map=None)
    if type(p).__name__ == "ConstantScenarioParameter":
        cs = p
        return dict(values=np.asarray(cs._values).reshape(1, -1).copy(), axis=cs._scenario_index, offset=0, col_map=None)
    return None


def scenario_axis(Parameter p):
    """Scenario axis of the per-scenario profile classes (evaluated through their public ``value``)."""
    cdef ScenarioMonthlyProfileParameter a
    cdef ScenarioDailyProfileParameter b
    cdef ScenarioWeeklyProfileParameter c
    cdef ArrayIndexedScenarioMonthlyFactorsParameter d
    cdef ConstantScenarioIndexParameter e
    name = type(p).__name__
    if name == "ConstantScenarioIndexParameter":
        e = p
        return e._scenario_index
    if name == "ScenarioMonthlyProfileParameter":
        a = p
        return a._scenario_index
    if name == "ScenarioDailyProfileParameter":
        b = p
        return b._scenario_index
    if name == "ScenarioWeeklyProfileParameter":
        c = p
        return c._scenario_index
    if name == "ArrayIndexedScenarioMonthlyFactorsParameter":
        d = p
        return d._scenario_index
    return None


def threshold_predicate(AbstractThresholdParameter p):
    """``AbstractThresholdParameter.predicate`` (pywr/parameters/_thresholds.pxd:12; enum Predicates, _thresholds.pyx:5-10)"""
    return int(p.predicate)


def multiple_thresholds(p):
    """``thresholds`` of MultipleThresholdParameterIndexParameter / MultipleThresholdIndexParameter (_thresholds.pxd:24,29)"""
    if isinstance(p, MultipleThresholdIndexParameter):
        return list((<MultipleThresholdIndexParameter>p).thresholds)
    return list((<MultipleThresholdParameterIndexParameter?>p).thresholds)


def polynomial1d_info(Polynomial1DParameter p):
    """the non-public members of Polynomial1DParameter (pywr/parameters/_polynomial.pxd:5-12)"""
    return dict(coefficients=np.asarray(p.coefficients).copy(), node=p._other_node, storage_node=p._storage_node,
                parameter=p._parameter, use_proportional_volume=bool(p.use_proportional_volume), offset=float(p.offset),
                scale=float(p.scale))


def polynomial2d_info(Polynomial2DStorageParameter p):
    """the non-public members of Polynomial2DStorageParameter (pywr/parameters/_polynomial.pxd:14-23)"""
    return dict(coefficients=np.asarray(p.coefficients).copy(), storage_node=p._storage_node, parameter=p._parameter,
                use_proportional_volume=bool(p.use_proportional_volume), storage_offset=float(p.storage_offset),
                storage_scale=float(p.storage_scale), parameter_offset=float(p.parameter_offset), parameter_scale=float(p.parameter_scale))


def threshold_values(AbstractThresholdParameter p):
    """``AbstractThresholdParameter.values`` (the pair the index selects from), or None"""
    if p.values is None:
        return None
    return np.asarray(p.values).copy()


def set_node_flow(AbstractNode node, double[:] flow):
    """node._flow[:] = flow ; node._prev_flow[:] = flow  (state after the last AbstractNode.after)"""
    cdef Py_ssize_t i
    for i in range(node._flow.shape[0]):
        node._flow[i] = flow[i]
        node._prev_flow[i] = flow[i]


def rolling_initial_utilisation(RollingVirtualStorage node):
    """``RollingVirtualStorage._initial_utilisation`` (pywr/_core.pyx:1500-1514): what the ring buffer holds at reset."""
    return node._initial_utilisation


def set_storage_state(AbstractStorage node, double[:] volume, double[:] pc):
    cdef Py_ssize_t i
    for i in range(node._volume.shape[0]):
        node._volume[i] = volume[i]
        node._current_pc[i] = pc[i]


def set_array_recorder(rec, double[:, :] data):
    """rec._data[:, :] = data for the NumpyArray node / storage recorder families."""
    cdef NumpyArrayNodeRecorder nr
    cdef NumpyArrayAbstractStorageRecorder sr
    cdef NumpyArrayParameterRecorder pr
    cdef NumpyArrayIndexParameterRecorder ir
    if isinstance(rec, NumpyArrayNodeRecorder):
        nr = rec
        nr._data[:, :] = data
    elif isinstance(rec, NumpyArrayAbstractStorageRecorder):
        sr = rec
        sr._data[:, :] = data
    elif isinstance(rec, NumpyArrayParameterRecorder):
        pr = rec
        pr._data[:, :] = data
    elif isinstance(rec, NumpyArrayIndexParameterRecorder):
        ir = rec
        ir._data = np.asarray(data).astype(np.int32)
    else:
        raise TypeError(f"unsupported recorder {type(rec).__name__}")


def temporal_agg_func(rec):
    """The temporal aggregation function of a NumpyArray recorder (its `temporal_agg_func` property is write-only,
    pywr/recorders/_recorders.pyx:607-609,1114-1116): a name, or the user's callable."""
    cdef NumpyArrayNodeRecorder nr
    cdef NumpyArrayAbstractStorageRecorder sr
    if isinstance(rec, NumpyArrayNodeRecorder):
        nr = rec
        return nr._temporal_aggregator.func
    if isinstance(rec, NumpyArrayAbstractStorageRecorder):
        sr = rec
        return sr._temporal_aggregator.func
    if isinstance(rec, NumpyArrayParameterRecorder):
        return rec._temporal_aggregator.func
    return None


def replace_array_recorder(rec, double[:, :] data):
    """rec._data = data: the recorder's array becomes `data` itself (any number of rows).  Used when a device run keeps
    the [T, S] arrays on the device: `data` is then ONE row holding the recorder's temporal aggregate per scenario, of which
    `values()` = aggregate_2d(_data, axis=0) (pywr/recorders/_recorders.pyx:630-633) is that same row for the sum / mean /
    min / max aggregation functions."""
    cdef NumpyArrayNodeRecorder nr
    cdef NumpyArrayAbstractStorageRecorder sr
    if isinstance(rec, NumpyArrayNodeRecorder):
        nr = rec
        nr._data = data
    elif isinstance(rec, NumpyArrayAbstractStorageRecorder):
        sr = rec
        sr._data = data
    elif isinstance(rec, NumpyArrayParameterRecorder):
        (<NumpyArrayParameterRecorder>rec)._data = data
    else:
        raise TypeError(f"unsupported recorder {type(rec).__name__}")


def set_constant_recorder(rec, double[:] values):
    """rec._values[:] = values for the one-value-per-scenario recorder families."""
    cdef BaseConstantNodeRecorder nr
    cdef BaseConstantStorageRecorder sr
    if isinstance(rec, BaseConstantNodeRecorder):
        nr = rec
        nr._values[:] = values
    elif isinstance(rec, BaseConstantStorageRecorder):
        sr = rec
        sr._values[:] = values
    elif isinstance(rec, BaseConstantParameterRecorder):
        (<BaseConstantParameterRecorder>rec)._values[:] = values
    else:
        raise TypeError(f"unsupported recorder {type(rec).__name__}")


def recorder_node(rec):
    cdef NodeRecorder nr
    cdef StorageRecorder sr
    if isinstance(rec, NodeRecorder):
        nr = rec
        return nr._node
    if isinstance(rec, StorageRecorder):
        sr = rec
        return sr._node
    return None
