import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
# the reference framework built by baseline/build_ref.sh (git-ignored, travels to the GPU box)
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(_REF, "pywr")) and _REF not in sys.path:
    sys.path.insert(0, _REF)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
