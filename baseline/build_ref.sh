#!/usr/bin/env bash
# Builds the UNMODIFIED reference (pywr, Cython framework only: libglpk / lpsolve55 are not in
# this image, so its two LP back-ends cannot be compiled) into baseline/_ref (git-ignored,
# travels to the GPU box).  Used as the *caller* of the drop-in solver (Model.run) and to run the
# reference's own known-answer tests against the oracle.  Nothing here is copied into the repo.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${PYWR_REFERENCE:-/root/reference}"
[ -d "$SRC/pywr" ] || { echo "reference not found at $SRC" >&2; exit 1; }
TMP="$(mktemp -d /tmp/pywr_build.XXXXXX)"
cp -r "$SRC/." "$TMP/"
# setuptools_scm is not installed: provide the version file it would have written
printf 'version = "1.31.1"\n__version__ = version\nversion_tuple = (1, 31, 1)\n' > "$TMP/pywr/_version.py"
# drop the setuptools-scm build requirement (not installable offline)
python - "$TMP/pyproject.toml" <<'PY'
import sys,re
p=sys.argv[1]; s=open(p).read()
s=s.replace('"setuptools-scm>=8", ','')
s=s.replace('dynamic = ["version"]','version = "1.31.1"')
s=re.sub(r'\[tool\.setuptools_scm\]\nversion_file = "pywr/_version.py"\n','',s)
s=s.replace('"pandas<3"','"pandas"')
open(p,'w').write(s)
PY
rm -rf "$HERE/_ref"
cd "$TMP"
PYWR_BUILD_GLPK=false PYWR_BUILD_LPSOLVE=false PACKAGE_DATA=false \
  python -m pip install --no-index --no-build-isolation --no-deps \
  --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP" 2>&1 | tail -5
# optional dependency stub so `import pywr` works without PyTables
mkdir -p "$HERE/_ref/tables"
cat > "$HERE/_ref/tables/__init__.py" <<'PY'
"""Stub: PyTables is not installed in this image; pywr only needs the import to succeed."""
class Filters:  # noqa
    def __init__(self, *a, **k):
        raise ImportError("PyTables stub")
def open_file(*a, **k):
    raise ImportError("PyTables stub")
PY
rm -rf "$TMP"
# The reference's own test-suite (test modules + model documents), staged next to the install so that it can also
# be run against the CUDA engine on the GPU box (oracle/run_reference_tests.sh b200).  baseline/_ref/ is
# git-ignored: nothing of the reference enters the history.
rm -rf "$HERE/_ref/_reftests"
cp -r "$SRC/tests" "$HERE/_ref/_reftests"
echo "built pywr into $HERE/_ref"
