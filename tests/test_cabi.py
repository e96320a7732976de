"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/b200lp.h declares; without a CUDA device it fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from pywr_b200.engine import load_library

    return load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "b200lp.h")).read()
    declared = set(re.findall(r"\b(b200lp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    from pywr_b200.engine import SYMBOLS

    bound = {name for name, _, _ in SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib, name)
    assert lib.b200lp_abi_version() == 8


def test_structure_struct_layout_matches_header():
    """The ctypes mirror must have one field per member of struct b200lp_structure, in order."""
    from pywr_b200.structure import b200lp_structure

    hdr = open(os.path.join(ROOT, "include", "b200lp.h")).read()
    start = hdr.index("typedef struct b200lp_structure {") + len("typedef struct b200lp_structure {")
    body = hdr[start: hdr.index("} b200lp_structure;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl or decl.startswith("typedef"):
            continue
        for part in decl.split(","):
            m = re.search(r"\*?\s*([a-z_0-9]+)\s*$", part.strip())
            names.append(m.group(1))
    assert names == [f for f, _ in b200lp_structure._fields_]


def test_no_cpu_fallback(lib):
    """Without a device create() must fail with B200LP_E_NO_DEVICE (and the Python engine raises)."""
    if lib.b200lp_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from fixtures_util import load_case
    from pywr_b200.engine import CudaEngine, EngineUnavailable

    s, _ = load_case("thames.route")
    with pytest.raises(EngineUnavailable):
        CudaEngine(s, 4)
    cs = s.to_c()
    h = ctypes.c_void_p()
    assert lib.b200lp_create(ctypes.addressof(cs), 4, 0, ctypes.byref(h)) == -2
    assert b"no CPU fallback" in lib.b200lp_last_error()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pywr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".pyx")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "lp_oracle" not in src.replace("oracle/lp_oracle.c", ""), f
    # the benchmark scripts measure the product only (bench.py's cpu_baseline / --impl reference legs are the one exception)
    for f in os.listdir(os.path.join(ROOT, "benchmarks")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "benchmarks", f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_structure_roundtrip(tmp_path):
    from fixtures_util import load_case
    from pywr_b200.structure import LPStructure

    s, _ = load_case("two_reservoir.route")
    assert (s.n_rows, s.n_cols, s.nnz, s.n_nodes) == (28, 13, 49, 26)
    p = tmp_path / "s.npz"
    s.save(p)
    s2 = LPStructure.load(p)
    assert np.array_equal(s.dense(), s2.dense())
    t, _ = load_case("thames.route")
    assert (t.n_rows, t.n_cols, t.nnz, t.n_nodes) == (15, 6, 23, 14)


def test_specialised_run_kernel_compiles_for_every_golden_program():
    """No device needed: b200lp_jit_probe generates the structure-specialised run kernel of every golden program and
    compiles it with NVRTC for sm_100a (the compile b200lp_load_program does on the GPU box); no register spills."""
    import ctypes

    from fixtures_util import load_program_case, n_scenarios_of, program_cases
    from pywr_b200.engine import load_library

    L = load_library()
    try:
        ctypes.CDLL("libnvrtc.so.12")
    except OSError:
        pytest.skip("libnvrtc is not installed here")
    for name in program_cases():
        s, prog, _, _ = load_program_case(name)
        cs, cp = s.to_c(), prog.to_c()
        log = ctypes.create_string_buffer(1 << 16)
        rc = L.b200lp_jit_probe(ctypes.addressof(cs), ctypes.addressof(cp), n_scenarios_of(s, prog), 1, log, len(log))
        assert rc == 0, (name, L.b200lp_last_error().decode())
        text = log.value.decode()
        assert "lp_run_jit" in text and " 0 bytes spill stores" in text.split("lp_run_jit")[-1], (name, text[-400:])
