#!/usr/bin/env python
"""Run on the GPU box: regenerate a small `database.csv` with the engine (main_compare.make_database)
for three geometries -- main_compare.py's own (r_outer = 0.037, pipe_offset = 0.0038) and two neighbours -- and
store it gzip-compressed where tests/test_reference_main_compare.py expects it.

    python tools/make_database_fixture.py [out.csv.gz]      # default gpurun_out/database_engine.csv.gz
"""
import gzip
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ray_tracing_ultrasound_b200 import main_compare  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "database_engine.csv.gz")
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "database.csv")
    n = main_compare.make_database(path, np.array([0.03, 0.037, 0.04]), np.array([0.004, 0.0038, 0.0]))
    with open(path, "rb") as f, gzip.open(out, "wb", compresslevel=9) as g:
        shutil.copyfileobj(f, g)
print(f"{n} rows -> {out} ({os.path.getsize(out)} bytes)")
