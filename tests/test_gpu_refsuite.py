"""The reference's OWN known-answer test-suite run against the CUDA engine (solver "b200") on the GPU:
every test that can run in this image (no PyTables / openpyxl / platypus, pandas 3) must pass; the set of
failing tests must stay within the environmental ones recorded in profiles/refsuite/b200_failed_tests.txt
(identical to the CPU oracle's).  Needs the suite staged by baseline/build_ref.sh (baseline/_ref/_reftests)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_suite_passes_on_the_cuda_engine():
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "_reftests")) and not os.path.isdir("/root/reference/tests"):
        pytest.skip("reference test-suite not staged (baseline/build_ref.sh)")
    pytest.importorskip("pywr")
    out = subprocess.run(["bash", os.path.join(ROOT, "oracle", "run_reference_tests.sh"), "b200", "--timeout", "300"],
                         capture_output=True, text=True, timeout=1500).stdout
    failed = {ln.split(" - ")[0] for ln in out.splitlines() if ln.startswith(("FAILED", "ERROR"))}
    known = {ln.strip() for ln in open(os.path.join(ROOT, "profiles", "refsuite", "b200_failed_tests.txt")) if ln.strip()}
    m = re.search(r"(\d+) passed", out)
    assert m, out[-2000:]
    assert failed <= known, sorted(failed - known)
    assert int(m.group(1)) >= 590
