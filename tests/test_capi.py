"""CPU: the C-ABI library loads and exports every symbol include/tofrt.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tofrt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tofrt_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from ray_tracing_ultrasound_b200 import _build, _capi
    if _build.is_stale():
        _build.build()
    lib = ctypes.CDLL(_build.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tofrt.h but not exported"
    assert sorted(_capi.EXPORTS) == declared
    lib.tofrt_version.restype = ctypes.c_int
    assert lib.tofrt_version() == 100


def test_no_cpu_fallback_without_gpu():
    """The product must fail loudly when there is no CUDA device / extension."""
    from ray_tracing_ultrasound_b200 import TofEngine, _capi
    lib = _capi.load()
    if lib.tofrt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        TofEngine(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or load it.
    (csrc/tofrt_math.h is shared the other way round: the oracle's twin build includes IT.)"""
    pkg = os.path.join(ROOT, "ray_tracing_ultrasound_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|oracle/|tofo_|#include\s+\"[^\"]*oracle)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # comments may mention it
                assert not pat.search(src), f"{f} references the oracle"


def test_reference_setup_matches_golden_inputs(golden):
    """x_a and alpha are rebuilt with the reference's NumPy expressions, bit for bit."""
    import numpy as np
    from ray_tracing_ultrasound_b200 import reference_setup as rs
    x_a, z_a = rs.transducer()
    assert np.array_equal(x_a, golden["sweep"]["x_a"])
    assert np.array_equal(rs.launch_angles(), golden["sweep"]["alpha"])
    assert np.array_equal(rs.launch_angles(rs.NUM_ALPHA_POINTS_COMPARE), golden["compare_live"]["alpha"])
    r, o = rs.sweep_grid()
    assert np.array_equal(r, golden["sweep"]["radius"]) and np.array_equal(o, golden["sweep"]["offset"])
    assert z_a[0] == 0.21 and rs.launch_angles()[452] == -1.1102230246251565e-16


def build_c_example(tmp_path):
    """examples/sweep.c against include/tofrt.h and libtofrt.so with a plain C99 compiler: the header is C, the
    boundary needs nothing but pointers and sizes."""
    from ray_tracing_ultrasound_b200 import _build
    if _build.is_stale():
        _build.build()
    exe = str(tmp_path / "sweep_example")
    libdir = os.path.dirname(_build.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "sweep.c"), "-L" + libdir, "-ltofrt", "-lm", "-o", exe], check=True)
    return exe, dict(os.environ, LD_LIBRARY_PATH=libdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))


def test_c_example_compiles_as_c99_and_fails_loudly_without_gpu(tmp_path):
    from ray_tracing_ultrasound_b200 import _capi
    exe, env = build_c_example(tmp_path)
    if _capi.load().tofrt_device_count() > 0:
        pytest.skip("a GPU is present (tests/test_gpu_dropin.py runs the example)")
    res = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
    assert res.returncode == 1 and "tofrt_create" in res.stderr and res.stdout == ""
