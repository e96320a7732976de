"""The reference's OWN known-answer test-suite with every ``Model.run()`` routed through the device-resident run (solver option
``device_run``; B200_REFSUITE_DEVICE_RUN in oracle/refsuite_plugin.py): the program compiler, the engines' program mode and the
write-back into Pywr's objects are pinned by the reference's answers for every parameter / recorder / node type the suite
exercises.  Models the device program does not cover fall back to the per-timestep loop (counted below).  CPU: the oracle's
program mode; GPU (-m gpu): the CUDA engine, with the generic and with the per-model specialised run kernel."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(solver, extra_env, select=None, min_passed=590, min_device=(290, 320)):
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "_reftests")) and not os.path.isdir("/root/reference/tests"):
        pytest.skip("reference test-suite not staged (baseline/build_ref.sh)")
    pytest.importorskip("pywr")
    env = dict(os.environ, B200_REFSUITE_DEVICE_RUN="1", **extra_env)
    cmd = ["bash", os.path.join(ROOT, "oracle", "run_reference_tests.sh"), solver, "--timeout", "300"]
    if select:      # a later -k replaces the script's own
        cmd += ["-k", f"not TestSpecifyingSolver and ({select})"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env).stdout
    failed = {ln.split(" - ")[0] for ln in out.splitlines() if ln.startswith(("FAILED", "ERROR"))}
    known = {ln.strip() for ln in open(os.path.join(ROOT, "profiles", "refsuite", "b200_failed_tests.txt")) if ln.strip()}
    m = re.search(r"(\d+) passed", out)
    d = re.search(r"device_run: (\d+) of (\d+) Model.run\(\) calls ran as device programs", out)
    assert m and d, out[-2000:]
    assert failed <= known, sorted(failed - known)
    assert int(m.group(1)) >= min_passed
    assert int(d.group(1)) >= min_device[0] and int(d.group(2)) >= min_device[1], d.group(0)   # 307 of 327 at the end of round 2


def test_reference_suite_through_the_device_run_of_the_oracle():
    _run("oracle", {})


@pytest.mark.gpu
def test_reference_suite_through_the_device_run_of_the_cuda_engine():
    """the whole suite on the generic run kernel (20 s)"""
    _run("b200", {"B200LP_JIT": "0"})


@pytest.mark.gpu
def test_reference_suite_through_the_specialised_run_kernel():
    """the kernel specialised per model costs one NVRTC compile per model of the suite (100 s for all of it: run by hand with
    B200_REFSUITE_DEVICE_RUN=1 oracle/run_reference_tests.sh b200, profiles/refsuite/device_run_summary.txt); here the modules
    around control curves, licences, aggregated / piecewise nodes, delays, scenarios and the run loop (134 runs, 40 s)"""
    _run("b200", {"B200LP_JIT": "1"}, min_passed=200, min_device=(125, 130),
         select="test_control_curves or test_delay or test_virtual or test_aggregated or test_piecewise or test_run or test_scenarios "
                "or test_license or test_core")


def test_edge_formulation_through_the_device_run_fails_nothing_the_per_timestep_run_passes():
    """the same on the edge formulation (CythonGLPKEdgeSolver's LP): the set of failing reference tests stays within that of the
    per-timestep run with the same solver (profiles/refsuite/edge_failed_tests.txt)"""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "_reftests")) and not os.path.isdir("/root/reference/tests"):
        pytest.skip("reference test-suite not staged (baseline/build_ref.sh)")
    pytest.importorskip("pywr")
    env = dict(os.environ, B200_REFSUITE_DEVICE_RUN="1")
    out = subprocess.run(["bash", os.path.join(ROOT, "oracle", "run_reference_tests.sh"), "oracle-edge", "--timeout", "300"],
                         capture_output=True, text=True, timeout=1500, env=env).stdout
    failed = {ln.split(" - ")[0] for ln in out.splitlines() if ln.startswith(("FAILED", "ERROR"))}
    base = {ln.strip() for ln in open(os.path.join(ROOT, "profiles", "refsuite", "edge_failed_tests.txt")) if ln.strip()}
    m = re.search(r"(\d+) passed", out)
    assert m and int(m.group(1)) >= 585, out[-1500:]
    assert failed <= base, sorted(failed - base)


@pytest.mark.gpu
def test_reference_suite_through_the_large_lp_path():
    """every model of the suite forced onto the large-LP path (B200LP_FORCE_PDHG + B200LP_FORCE_RSX): presolve with its equality-row
    eliminations, revised simplex with the basis inverse in HBM, plane kernels around it, as device programs.  Fails nothing the
    per-timestep edge run of the oracle passes (profiles/refsuite/edge_failed_tests.txt), except the model with dynamic aggregated factors (a4: rejected on this path)."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "_reftests")) and not os.path.isdir("/root/reference/tests"):
        pytest.skip("reference test-suite not staged (baseline/build_ref.sh)")
    pytest.importorskip("pywr")
    script = os.path.join(ROOT, "oracle", "run_reference_tests.sh")
    env = dict(os.environ, B200_REFSUITE_DEVICE_RUN="1", B200LP_FORCE_PDHG="1", B200LP_FORCE_RSX="1")
    out = subprocess.run(["bash", script, "b200-edge", "--timeout", "300"], capture_output=True, text=True, timeout=1500, env=env).stdout
    # what the per-timestep edge run of the oracle fails in this image (oracle/run_reference_tests.sh oracle-edge)
    base = {ln.strip() for ln in open(os.path.join(ROOT, "profiles", "refsuite", "edge_failed_tests.txt")) if ln.strip()}
    extra = {ln.split(" - ")[0] for ln in out.splitlines() if ln.startswith(("FAILED", "ERROR"))} - base
    assert extra <= {"FAILED test_run.py::test_loss_link_node_loss_factor_parameter"}, sorted(extra)
    m = re.search(r"(\d+) passed", out)
    d = re.search(r"device_run: (\d+) of (\d+)", out)
    assert m and int(m.group(1)) >= 584 and d and int(d.group(1)) >= 295, out[-1500:]


@pytest.mark.gpu
def test_reference_suite_subset_through_pdhg_and_crossover():
    """the first-order method of the large-LP path followed by the crossover to a vertex (B200LP_LARGE_ENGINE=pdhg) on the modules
    around the analytical reservoir model, delays, piecewise and aggregated nodes, licences: scenarios the iterations leave at their
    limit are finished by the simplex of the crossover (test_analytical.py::test_analytical needs that).  The whole suite this way:
    586 passed, 50 s (profiles/refsuite/device_run_summary.txt)."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "_reftests")) and not os.path.isdir("/root/reference/tests"):
        pytest.skip("reference test-suite not staged (baseline/build_ref.sh)")
    pytest.importorskip("pywr")
    env = dict(os.environ, B200_REFSUITE_DEVICE_RUN="1", B200LP_FORCE_PDHG="1", B200LP_LARGE_ENGINE="pdhg")
    out = subprocess.run(["bash", os.path.join(ROOT, "oracle", "run_reference_tests.sh"), "b200-edge", "--timeout", "200", "-k",
                          "not TestSpecifyingSolver and (test_analytical or test_delay or test_piecewise or test_license or test_aggregated)"],
                         capture_output=True, text=True, timeout=1200, env=env).stdout
    base = {ln.strip() for ln in open(os.path.join(ROOT, "profiles", "refsuite", "edge_failed_tests.txt")) if ln.strip()}
    extra = {ln.split(" - ")[0] for ln in out.splitlines() if ln.startswith(("FAILED", "ERROR"))} - base
    assert not extra, sorted(extra)
    m = re.search(r"(\d+) passed", out)
    assert m and int(m.group(1)) >= 40, out[-1500:]
