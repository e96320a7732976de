"""Compiles pywr_b200/bridge/_bridge.pyx in place against the reference build found on sys.path
(baseline/_ref by default).  Called by __graft_entry__.build()."""
import os
import subprocess
import sys
import sysconfig


def build(ref_dir):
    here = os.path.dirname(os.path.abspath(__file__))
    pyx = os.path.join(here, "_bridge.pyx")
    c_file = os.path.join(here, "_bridge.c")
    so = os.path.join(here, "_bridge" + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(pyx):
        return so
    import numpy

    subprocess.check_call([sys.executable, "-m", "cython", "-3", "-I", ref_dir, pyx, "-o", c_file])
    inc = sysconfig.get_paths()["include"]
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                           "-I", inc, "-I", numpy.get_include(), c_file, "-o", so])
    os.remove(c_file)
    return so


if __name__ == "__main__":
    print(build(sys.argv[1]))
