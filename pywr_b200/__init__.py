"""pywr_b200 — B200-native scenario-batched LP engine behind Pywr's solver API.

Scope: the one hot path of pywr/pywr named by BASELINE.json — the per-timestep, per-scenario loop
of ``CythonGLPKSolver.solve`` (``pywr/solvers/cython_glpk.pyx:821-1095``) with flow commit, storage
mass balance and the built-in array recorders — as hand-written sm_100a CUDA behind the C-ABI in
``include/b200lp.h``.  Importing this package registers the solvers ``b200`` (route formulation)
and ``b200-edge`` (edge formulation) with Pywr when Pywr is importable.
"""

__version__ = "0.1.0"


def register():
    from .solver import register as _register

    return _register()


try:  # registration is a no-op when pywr is not installed (C-ABI-only use)
    import pywr.solvers  # noqa: F401
except Exception:  # pragma: no cover
    pass
else:
    register()
