import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz"))
            for name in ("database_2", "compare", "sweep", "rays", "compare_live", "shooters")}


@pytest.fixture(scope="session")
def golden_array128():
    """Live reference with a 128-element array (tests/golden/make_golden.py array128)."""
    return np.load(os.path.join(GOLDEN, "array128.npz"))


@pytest.fixture(scope="session")
def oracle_libm():
    from oracle.oracle import Oracle
    return Oracle(device_math=False)


@pytest.fixture(scope="session")
def oracle_dm():
    from oracle.oracle import Oracle
    return Oracle(device_math=True)


@pytest.fixture(scope="session")
def engine():
    """The product, through the C ABI.  Builds libtofrt.so if the source is newer."""
    from ray_tracing_ultrasound_b200 import TofEngine, _build, _capi
    if _build.is_stale():
        _build.build()
    if _capi.load().tofrt_device_count() == 0:
        pytest.skip("no CUDA device: the engine has no CPU fallback (run `pytest -m gpu` on the B200 box)")
    return TofEngine(0)


@pytest.fixture(params=["angle", "vector"])
def math_mode(request, engine, oracle_dm):
    """Runs a GPU test once per TOFRT_OPT_MATH mode, with the oracle twin switched alike
    (modules opt in with pytest.mark.usefixtures("math_mode"))."""
    default = engine.math_mode
    engine.set_math(request.param)
    oracle_dm.set_math(request.param == "vector")
    yield request.param
    engine.set_math(default)
    oracle_dm.set_math(False)
