"""CPU, build container only: the oracle against the LIVE, unmodified reference on random
geometries beyond the committed fixtures.  Skipped where /root/reference does not exist
(the GPU box); the committed goldens (tests/golden) cover that case."""
import os
import sys
import warnings

import numpy as np
import pytest

REF = "/root/reference/main_rt.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="live reference not mounted")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    warnings.simplefilter("ignore")
    return make_golden.load_reference(), make_golden


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_matches_live_reference_on_random_geometries(ref, oracle_libm, oracle_dm, seed):
    r, mg = ref
    rng = np.random.default_rng(100 + seed)
    x_a, z_a, zf, alpha = mg.array_and_alpha(r, 181 * 5)
    for _ in range(5):
        rad, off = rng.uniform(0.02, 0.1), rng.uniform(-0.01, 0.01)
        shooter = 32 if seed == 0 else int(rng.integers(20, 45))
        r.r_outer, r.pipe_offset = np.float64(rad), np.float64(off)
        try:
            res = r.shoot_rays(x_a[shooter], z_a[shooter], zf, alpha, plot=False)
        except np.linalg.LinAlgError:
            continue  # a ray missed the pipe: outside the reference's domain
        hit, k_hit, tof = mg.assemble_table(r, x_a, z_a, shooter, res, atol=1e-6)
        ridx, margin, _, _, _ = mg.reference_idx(r, x_a[shooter], z_a[shooter], alpha, res)
        for o in (oracle_libm, oracle_dm):
            p = o.params()
            t, k, idx = o.table(p, [rad], [off], x_a[shooter], z_a[shooter], zf, alpha, x_a, want_idx=True)
            assert np.array_equal(k[:, 0], k_hit)
            assert np.abs(t[:, 0] - tof).max() < 1e-15
            oi = np.where(idx[0] < -1, -1, idx[0])
            bad = (oi != ridx) & (ridx != -2)
            assert not (bad & (margin > 1e-13)).any()
            got = o.shoot_rays(p, rad, off, x_a[shooter], z_a[shooter], zf, alpha)
            for key in ("pipe_x", "pipe_z", "lens_2_x", "lens_2_z"):
                assert np.array_equal(np.isnan(got[key]), np.isnan(res[key]))
                assert np.nanmax(np.abs(got[key] - res[key])) < 1e-13
