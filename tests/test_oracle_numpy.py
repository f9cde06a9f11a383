"""CPU: the vectorised NumPy restatement (oracle/oracle_numpy.py, "oracle-A": what bench.py times as the reference's
NumPy CPU path) against the reference's fixtures and the live-reference goldens."""
import numpy as np

from oracle import oracle_numpy as on
from ray_tracing_ultrasound_b200 import reference_setup as rs


def _setup(n_el=64, K=905):
    x_a, z_a = rs.transducer(num_elements=n_el)
    return on.Setup(x_a, z_a, rs.focus_line(K), rs.launch_angles(K), n_el // 2)


def test_numpy_port_reproduces_database_2_and_the_live_sweep(golden):
    S = _setup()
    radii, offsets = rs.sweep_grid()
    sel = np.arange(0, 210, 3)                       # every third geometry keeps the CPU suite short
    tof, k_hit = on.table(S, radii[sel], offsets[sel])
    db, sw = golden["database_2"], golden["sweep"]
    assert np.array_equal(k_hit >= 0, db["hitted"].reshape(210, 65)[sel])
    assert np.abs(tof - db["tof_total"].reshape(210, 65)[sel]).max() <= 1e-12
    assert np.array_equal(k_hit, sw["k_hit"][sel])
    assert np.abs(tof - sw["tof"][sel]).max() <= 1e-12
    # per-ray: idx equals the verbatim reference's wherever its decision is not rounding noise; same NaN pattern
    for g in (1, 57, 100, 209):
        res = on.shoot_rays(S, radii[g], offsets[g])
        ridx = sw["idx"][g].astype(np.int32)          # -1 no sign change, -2 NaN line
        mine = np.where(res["idx"] < -1, -1, res["idx"])
        differs = (mine != ridx) & (ridx != -2)
        assert not (differs & (sw["margin"][g] > 1e-13)).any()
        nan_ref = np.unpackbits(sw["nan_target"][g])[:905].astype(bool)
        assert np.array_equal(np.isnan(res["target_x"]), nan_ref)


def test_numpy_port_matches_the_c_oracle_on_config_3_geometries(oracle_libm):
    S = _setup(128)
    radii, offsets = rs.synthetic_grid(1024, 1024)
    sel = np.linspace(0, len(radii) - 1, 24).astype(np.int64)
    tof, k_hit = on.table(S, radii[sel], offsets[sel])
    p = oracle_libm.params()
    t2, k2, _ = oracle_libm.table(p, radii[sel], offsets[sel], S.x_a[64], S.z_a[64], S.zf, S.alpha, S.x_a)
    assert np.array_equal(k_hit, k2.T) and np.abs(tof - t2.T).max() <= 1e-15
    assert (k_hit >= 0).sum() > 40
