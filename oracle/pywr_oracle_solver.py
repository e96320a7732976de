"""Pywr solver plug-in backed by the CPU ORACLE.  TEST INFRASTRUCTURE ONLY.

Registers solver ``oracle`` (route formulation) and ``oracle-edge`` with the built reference so that
the reference's own known-answer test-suite can be run against the oracle
(``PYWR_SOLVER=oracle pytest <reference>/tests``), which is how the oracle is pinned
(SURVEY.md section 8c).  It reuses the product's host-side logic (structure compiler, gather,
commit) and swaps only the engine.
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pywr_b200.solver import BatchedLPSolver  # noqa: E402

from oracle.oracle_engine import OracleEngine  # noqa: E402


class OracleSolver(BatchedLPSolver):
    name = "oracle"
    formulation = "route"

    def _make_engine(self, structure, n_scenarios):
        return OracleEngine(structure, n_scenarios, **self.engine_args)


class OracleEdgeSolver(OracleSolver):
    name = "oracle-edge"
    formulation = "edge"


def register(first=False):
    import pywr.solvers as ps

    for cls in (OracleSolver, OracleEdgeSolver):
        if not any(getattr(c, "name", None) == cls.name for c in ps.solver_registry):
            if first:
                ps.solver_registry.insert(0, cls)
            else:
                ps.solver_registry.append(cls)


register()


class OracleRsxEdgeSolver(OracleSolver):
    """Edge formulation solved by the HOST build of the revised-simplex core (oracle/rsx_oracle.py): fast enough for
    host runs of the 2,073-node network (tests/golden/make_network_program.py)."""
    name = "oracle-rsx-edge"
    formulation = "edge"

    def _make_engine(self, structure, n_scenarios):
        from oracle.rsx_oracle import RsxOracleEngine

        return RsxOracleEngine(structure, n_scenarios)


def _register_rsx():
    import pywr.solvers as ps

    if not any(getattr(c, "name", None) == OracleRsxEdgeSolver.name for c in ps.solver_registry):
        ps.solver_registry.append(OracleRsxEdgeSolver)


_register_rsx()
