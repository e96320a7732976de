"""pytest plugin (``-p oracle.refsuite_plugin``) that plugs a solver of this repository into the
built reference BEFORE the reference's test modules are collected, so that the reference's own
known-answer suite pins it.  TEST INFRASTRUCTURE ONLY.

``B200_REFSUITE_SOLVER`` selects the solver: ``oracle`` / ``oracle-edge`` (CPU oracle, default) or
``b200`` / ``b200-edge`` (the CUDA engine, GPU box only).
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (_ROOT, os.path.join(_ROOT, "baseline", "_ref")):
    if p not in sys.path:
        sys.path.insert(0, p)

import pywr.solvers as ps  # noqa: E402

name = os.environ.get("B200_REFSUITE_SOLVER", "oracle")
if name.startswith("oracle"):
    import oracle.pywr_oracle_solver  # noqa: E402,F401
else:
    import pywr_b200  # noqa: E402,F401
from pywr_b200.solver import B200LPError  # noqa: E402

# the reference suite reads pywr.solvers.GLPKError at import time (tests/test_run.py:651,1241);
# it only exists when the GLPK extension was built
if not hasattr(ps, "GLPKError"):
    ps.GLPKError = B200LPError
# put the selected solver first: Model() with no solver argument takes solver_registry[0]
cls = next(c for c in ps.solver_registry if c.name == name)
ps.solver_registry.remove(cls)
ps.solver_registry.insert(0, cls)
# Some reference tests only run when the solver is *called* "glpk"/"glpk-edge" (dynamic aggregated
# factors, tests/test_agg_constraints.py:332,376).  B200_REFSUITE_ALIAS=1 registers the solver under
# the reference's name so that those tests exercise it too.
if os.environ.get("B200_REFSUITE_ALIAS"):
    cls.name = "glpk-edge" if name.endswith("edge") else "glpk"
    name = cls.name
# B200_REFSUITE_DEVICE_RUN=1: every Model.run() of the suite goes through the device-resident run (solver option device_run,
# pywr_b200/solver.py) - the program compiler, the program mode of the engine and the write-back into Pywr's objects are then
# pinned by the reference's known answers too; models the device program does not cover fall back to the per-timestep loop
if os.environ.get("B200_REFSUITE_DEVICE_RUN"):
    _init = cls.__init__

    def _init_device_run(self, *args, **kwargs):
        kwargs.setdefault("device_run", True)
        _init(self, *args, **kwargs)

    cls.__init__ = _init_device_run
os.environ["PYWR_SOLVER"] = name


def pytest_terminal_summary(terminalreporter):
    """with B200_REFSUITE_DEVICE_RUN: how many Model.run() calls ran as device programs, and why the others did not"""
    if not os.environ.get("B200_REFSUITE_DEVICE_RUN"):
        return
    import collections

    from pywr_b200.solver import device_run_log

    n_dev = sum(1 for kind, _ in device_run_log if kind == "device")
    terminalreporter.write_line(f"device_run: {n_dev} of {len(device_run_log)} Model.run() calls ran as device programs")
    why = collections.Counter(reason.split(":")[0][:70] for kind, reason in device_run_log if kind == "host")
    for reason, count in why.most_common(25):
        terminalreporter.write_line(f"  host x{count}: {reason}")
