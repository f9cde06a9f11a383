"""GPU: the drop-in entry points (main_rt.py sweep -> database_2.csv, main_compare.py ->
compare.csv + MSE against database.csv) keep the reference's formats and numbers."""
import csv
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOF_TOL = 1e-12


def test_main_rt_writes_database_2_csv(tmp_path, golden, engine):
    from ray_tracing_ultrasound_b200 import main_rt
    main_rt._engine, main_rt._engine_key = engine, None
    out = tmp_path / "database_2.csv"
    out.write_bytes(b"elem_idx,offset,radius,hitted,tof_total\r\n")  # the reference's header (main_rt.py:484)
    main_rt.main(str(out))
    raw = out.read_bytes()
    assert raw.count(b"\r\n") == 13651 and b"\n" not in raw.replace(b"\r\n", b"")
    rows = list(csv.reader(out.read_text().splitlines()))
    db = golden["database_2"]
    assert rows[0] == ["elem_idx", "offset", "radius", "hitted", "tof_total"]
    body = rows[1:]
    assert len(body) == 13650
    assert [int(r[0]) for r in body] == db["elem_idx"].tolist()
    assert np.array_equal(np.array([float(r[1]) for r in body]), db["offset"])
    assert np.array_equal(np.array([float(r[2]) for r in body]), db["radius"])
    assert [r[3] == "True" for r in body] == db["hitted"].tolist()
    assert np.abs(np.array([float(r[4]) for r in body]) - db["tof_total"]).max() <= TOF_TOL
    # text format of the reference's rows (not-hit rows are byte-identical)
    text = [",".join(r) for r in rows]
    for i, want in zip(list(range(4)) + list(range(32, 36)), db["head_text"]):
        got = text[i]
        if ",True," in want:
            assert got.rsplit(",", 1)[0] == want.rsplit(",", 1)[0]
            assert abs(float(got.rsplit(",", 1)[1]) - float(want.rsplit(",", 1)[1])) <= TOF_TOL
        else:
            assert got == want
    # append semantics of main_rt.py:502: a second run appends another 13 650 rows
    main_rt.main(str(out))
    assert out.read_bytes().count(b"\r\n") == 27301


def test_shoot_rays_dropin_matches_live_reference(golden, engine):
    """Same call as main_rt.py:481-482 with the module globals set the reference's way."""
    from ray_tracing_ultrasound_b200 import main_rt, reference_setup as rs
    main_rt._engine, main_rt._engine_key = engine, None
    rays = golden["rays"]
    x_a, z_a = rs.transducer()
    zf, alpha = rs.focus_line(), rs.launch_angles()
    for i in (0, 5, 11):
        main_rt.r_outer, main_rt.pipe_offset = rays["radius"][i], rays["offset"][i]
        res = main_rt.shoot_rays(x_a[32], z_a[32], zf, alpha, plot=False)
        assert sorted(res) == ["lens_1_x", "lens_1_z", "lens_2_x", "lens_2_z", "pipe_x", "pipe_z", "target_x", "target_z"]
        assert all(v.dtype == np.float64 and v.shape == (905,) for v in res.values())
        assert res["target_z"] is not zf and np.array_equal(res["target_z"], zf)
        # Non-degenerate geometries: every point within a few ulp of the live reference.  In a degenerate geometry
        # r = |offset| the pipe touches the origin, where every ray of the constant-TOF lens is aimed: all rays hit
        # the pipe at its edge point, z_q = sqrt(r^2 - (x_q - offset)^2) amplifies the last bit of x_q to ~1e-13 there,
        # and the centre ray (x_q ~ 1e-17) has z_q ~ 1e-10 with a relative error of order one in EVERY formulation --
        # the reference's own z_q = a_pq x_q + b_pq cancels 16 digits there.
        degenerate = abs(abs(float(main_rt.pipe_offset)) - float(main_rt.r_outer)) < 1e-15
        for key, tol in (("pipe_x", 1e-15), ("pipe_z", 1e-15), ("lens_2_x", 1e-14), ("lens_2_z", 1e-14)):
            a, b = res[key], rays[key][i]
            assert np.array_equal(np.isnan(a), np.isnan(b))
            d = np.nan_to_num(np.abs(a - b))
            if degenerate:
                assert (d > 1e-12).sum() <= 1 and d.max() <= 1e-9, (key, i)
            else:
                assert d.max() <= tol, (key, i, float(d.max()))
        a, b = res["target_x"], rays["target_x"][i]
        ok = ~np.isnan(b)
        assert np.array_equal(np.isnan(a), ~ok)
        assert (np.abs(a[ok] - b[ok]) <= 1e-11 * np.maximum(1.0, np.abs(b[ok]))).all()
    with pytest.raises(ValueError):
        main_rt.shoot_rays(0.0, 0.21, np.array([0.21]), np.array([0.0]), plot=False)
    with pytest.raises(ValueError):
        main_rt.shoot_rays(0.0, 0.21, np.zeros(3), np.zeros(4), plot=False)


def test_main_compare_roundtrip(tmp_path, golden, engine, capsys):
    """database.csv is absent from the reference checkout: regenerate it with the engine, then
    main_compare's MSE against it must be exactly 0 and compare.csv must match the fixture."""
    pd = pytest.importorskip("pandas")
    from ray_tracing_ultrasound_b200 import main_compare
    db_path, cmp_path = tmp_path / "database.csv", tmp_path / "compare.csv"
    radii = np.array([0.03, 0.037, 0.04])
    offsets = np.array([0.004, 0.0038, 0.0])
    n = main_compare.make_database(str(db_path), radii, offsets, engine=engine)
    assert n == 3 * 1810
    errs, num_hitted, mse = main_compare.main(str(db_path), str(cmp_path), engine=engine)
    printed = capsys.readouterr().out
    assert "num_hitted = " in printed and "mse = " in printed
    assert int(num_hitted) == 393 and float(mse) == 0.0
    assert np.all(errs[~np.isnan(errs)] == 0.0) and len(errs) == 1810
    raw = cmp_path.read_bytes()
    assert raw.startswith(b"alpha,offset,radius,hitted,tof_1,tof_2,tof_3,tof_4\r\n") and raw.count(b"\r\n") == 1811
    cc = golden["compare"]
    df = pd.read_csv(cmp_path, float_precision="round_trip")  # the default parser is off by an ulp
    assert np.array_equal(df["alpha"].to_numpy(), cc["alpha"])
    assert np.array_equal(df["hitted"].to_numpy(), cc["hitted"])
    for i in range(4):
        a, b = df[f"tof_{i + 1}"].to_numpy(), cc["tof"][i]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.nanmax(np.abs(a - b)) <= TOF_TOL
    # first data row of the reference's compare.csv: everything but the last float digits is text-identical
    first = raw.split(b"\r\n")[1].decode().split(",")
    want = str(cc["head_text"][1]).split(",")
    assert first[:4] == want[:4] and first[7] == want[7] == "nan"


def test_c_example_runs_the_reference_sweep_through_the_c_abi(tmp_path, golden, engine):
    """examples/sweep.c: main_rt.py's sweep from plain C (tofrt_setup + one tofrt_table_host call), its CSV against
    the reference's database_2.csv fixture: same rows, same hitted flags, travel times within 1e-12 s."""
    import subprocess
    from test_capi import build_c_example
    exe, env = build_c_example(tmp_path)
    out = tmp_path / "database_c.csv"
    res = subprocess.run([exe, str(out)], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert res.stdout.split() == "geometries 210 receivers 65 rows 13650 hits 498".split()
    raw = out.read_bytes()
    assert raw.count(b"\r\n") == 13650
    rows = [r.split(",") for r in raw.decode().split("\r\n") if r]
    db = golden["database_2"]
    assert [int(r[0]) for r in rows] == db["elem_idx"].tolist()
    assert np.abs(np.array([float(r[1]) for r in rows]) - db["offset"]).max() < 1e-17
    assert [r[3] == "True" for r in rows] == db["hitted"].tolist()
    assert np.abs(np.array([float(r[4]) for r in rows]) - db["tof_total"]).max() <= TOF_TOL
