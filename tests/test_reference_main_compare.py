"""CPU, build container only: the reference's OWN main_compare.py, UNMODIFIED, against a database.csv the engine
produced ("then main_compare.py runs unmodified").

tests/golden/database_engine.csv.gz was written on the GPU box by tools/make_database_fixture.py
(ray_tracing_ultrasound_b200.main_compare.make_database: three geometries x 1 810 rays, the per-ray records of the
CUDA path).  Here /root/reference/main_compare.py is executed as it is -- cwd = a scratch directory holding that
database.csv, a no-op `matplotlib` stub on PYTHONPATH (matplotlib is not installed; the script only plots with it) --
and its own printout is checked: num_hitted = 393 (as in the reference's compare.csv) and an MSE of its freshly
traced NumPy travel times against the engine's database below 1e-30 s^2.  Skipped where /root/reference does not
exist (the GPU box)."""
import gzip
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN

REF = "/root/reference/main_compare.py"
FIXTURE = os.path.join(GOLDEN, "database_engine.csv.gz")

pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="the reference checkout is not mounted here")


def test_unmodified_reference_main_compare_against_the_engines_database(tmp_path):
    pytest.importorskip("pandas")
    with gzip.open(FIXTURE, "rb") as g, open(tmp_path / "database.csv", "wb") as f:
        shutil.copyfileobj(g, f)
    stub = tmp_path / "stub" / "matplotlib"
    stub.mkdir(parents=True)
    (stub / "__init__.py").write_text("")
    (stub / "pyplot.py").write_text("def __getattr__(name):\n    return lambda *a, **k: None\n")
    env = dict(os.environ, PYTHONPATH=str(tmp_path / "stub"), MPLBACKEND="Agg")
    res = subprocess.run([sys.executable, REF], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    out = res.stdout
    m_hit = re.search(r"num_hitted = (?:np\.int64\()?(\d+)", out)
    m_mse = re.search(r"mse = (?:np\.float64\()?([0-9.eE+-]+|nan)", out)
    assert m_hit and m_mse, out[-500:]
    assert int(m_hit.group(1)) == 393
    mse = float(m_mse.group(1))
    assert mse < 1e-30, mse
    # the script rewrote compare.csv from its own NumPy trace: same hit flags as the reference's fixture
    import pandas as pd
    df = pd.read_csv(tmp_path / "compare.csv")
    cc = np.load(os.path.join(GOLDEN, "compare.npz"))
    assert len(df) == 1810 and np.array_equal(df["hitted"].to_numpy(), cc["hitted"])
    # and the engine's database rows for that geometry carry the same hit flags and legs within 1e-12 s
    db = pd.read_csv(tmp_path / "database.csv", float_precision="round_trip")
    sl = db[(db["offset"] == 0.0038) & (db["radius"] == 0.037)]
    assert len(sl) == 1810 and np.array_equal(sl["hitted"].to_numpy(), cc["hitted"])
    for i in range(4):
        a, b = sl[f"tof_{i + 1}"].to_numpy(), cc["tof"][i]
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.nanmax(np.abs(a - b)) <= 1e-12
