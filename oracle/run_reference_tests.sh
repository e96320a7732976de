#!/usr/bin/env bash
# Runs the reference's own test-suite (read-only under /root/reference/tests) against a solver of
# this repository.  usage: oracle/run_reference_tests.sh [oracle|oracle-edge|b200|b200-edge] [pytest args]
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SOLVER="${1:-oracle}"; shift || true
REF="${PYWR_REFERENCE:-/root/reference}"
TESTS="$REF/tests"
# GPU box: /root/reference does not exist there; baseline/build_ref.sh stages the suite next to the install
[ -d "$TESTS" ] || TESTS="$ROOT/baseline/_ref/_reftests"
WORK="$(mktemp -d /tmp/refsuite.XXXXXX)"
cd "$WORK"
B200_REFSUITE_SOLVER="$SOLVER" PYTHONPATH="$ROOT:$ROOT/baseline/_ref" \
  python -m pytest "$TESTS" -p oracle.refsuite_plugin -p no:cacheprovider -q --no-header \
  --rootdir "$WORK" -o addopts="" \
  --continue-on-collection-errors \
  --ignore "$TESTS/test_bisect.py" --ignore "$TESTS/test_notebook.py" \
  -k "not TestSpecifyingSolver" "$@"
