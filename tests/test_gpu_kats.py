"""GPU: the analytic known answers of SURVEY.md section 4, asked of the CUDA path itself (both math modes) -- what
the reference's own unit tests would be if it had any: constant-TOF lens, h(0) = h0, dh/dalpha against a finite
difference, the normal-incidence echo, Snell's law at both lens crossings and the mirror law at the pipe, all read off
the dict of shoot_rays (main_rt.py:432-441) that the kernel fills."""
import numpy as np
import pytest

from ray_tracing_ultrasound_b200 import reference_setup as rs
from fermat_oracle import Lens

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("math_mode")]


def _std(engine):
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    engine.setup(alpha, zf, x_a[32], z_a[32], x_a)
    return x_a, z_a, alpha


def test_lens_surface_known_answers(engine):
    """main_rt.py:181-234: every lens sample is T = l0/c1 + h0/c2 away from (array centre, origin); h(0) = h0; the
    analytic dh/dalpha the kernel uses (through the first refraction) matches a central difference of h."""
    _, _, alpha = _std(engine)
    lens = Lens()
    lx, lz = engine.lens()
    T = lens.l0 / lens.c1 + lens.h0 / lens.c2
    assert np.abs(np.hypot(lx, lz - lens.d) / lens.c1 + np.hypot(lx, lz) / lens.c2 - T).max() < 1e-18
    engine.setup(np.array([-1e-3, 0.0, 1e-3]), np.full(3, float(rs.D)), 0.0, float(rs.D), np.array([0.0]))
    x0, z0 = engine.lens()
    assert x0[1] == 0.0 and abs(z0[1] - lens.h0) < 1e-16
    eps = 1e-6
    fd = (lens.h(alpha + eps) - lens.h(alpha - eps)) / (2 * eps)
    assert np.abs(lens.dh(alpha) - fd).max() < 1e-9          # the oracle's analytic form ...
    _std(engine)
    # ... and the kernel's: direction of the refracted ray P_k -> Q_k obeys Snell's law with the normal built from dh
    rays = engine.shoot_rays(0.03, 0.0)
    _check_snell_first_crossing(lens, alpha, rays)


def _unit(vx, vz):
    n = np.hypot(vx, vz)
    return vx / n, vz / n


def _check_snell_first_crossing(lens, alpha, rays, tol=1e-12):
    """sin(theta_2)/sin(theta_1) = c2/c1 for A -> P_k (lens) -> Q_k (water), angles against the surface normal at P_k."""
    px, pz, qx, qz = rays["lens_1_x"], rays["lens_1_z"], rays["pipe_x"], rays["pipe_z"]
    ok = np.isfinite(qx)
    assert ok.mean() > 0.5
    tx, tz = _unit(*lens.tangent(alpha))                     # unit tangent; sin of the angle to the normal = |d . t|
    d1x, d1z = _unit(px - 0.0, pz - lens.d)
    d2x, d2z = _unit(qx - px, qz - pz)
    s1, s2 = d1x * tx + d1z * tz, d2x * tx + d2z * tz
    assert np.abs(s2 - (lens.c2 / lens.c1) * s1)[ok].max() < tol


@pytest.mark.parametrize("r", [0.01, 0.02, 0.03, 0.05])
def test_normal_incidence_echo(engine, r):
    """Centre ray, centred pipe, r < h0: down to the pipe apex and back, tof = 2 (l0/c1 + (h0 - r)/c2); the
    database_2.csv row (32, 0.0, 0.01) is this number."""
    _std(engine)
    lens = Lens()
    t = engine.tof_table([r], [0.0])
    assert t.k_hit[0, 32] >= 0
    assert abs(t.tof[0, 32] - 2 * (lens.l0 / lens.c1 + (lens.h0 - r) / lens.c2)) < 5e-18


@pytest.mark.parametrize("r,off", [(0.03, 0.0), (0.05, 0.004), (0.02, -0.007), (0.07, 0.001)])
def test_snell_and_mirror_laws_along_the_ray(engine, r, off):
    """Per ray, from the kernel's own points: Snell at P_k; the mirror law at Q_k about the reference's pipe normal
    (dzdx_pipe ignores the offset, main_rt.py:237-238: the normal at Q is (x_q, sqrt(r^2 - x_q^2)) whatever the
    offset -- the quirk is part of the contract); Snell back into the lens at the re-entry point with the analytic
    tangent at alpha' = atan2(x', z') (main_rt.py:396-398); and the landing point on the array line."""
    _, _, alpha = _std(engine)
    lens = Lens()
    rays = engine.shoot_rays(r, off)
    _check_snell_first_crossing(lens, alpha, rays)
    px, pz, qx, qz = rays["lens_1_x"], rays["lens_1_z"], rays["pipe_x"], rays["pipe_z"]
    ix, iz, tx_ = rays["lens_2_x"], rays["lens_2_z"], rays["target_x"]
    back = np.isfinite(ix)
    assert back.sum() > 100
    # mirror law: incident and reflected directions have opposite normal and equal tangential components
    nx, nz = _unit(qx, np.sqrt(r * r - qx * qx))
    dinx, dinz = _unit(qx - px, qz - pz)
    doutx, doutz = _unit(ix - qx, iz - qz)
    tang_in, tang_out = dinx * nz - dinz * nx, doutx * nz - doutz * nx
    norm_in, norm_out = dinx * nx + dinz * nz, doutx * nx + doutz * nz
    assert np.abs(tang_in - tang_out)[back].max() < 1e-9 and np.abs(norm_in + norm_out)[back].max() < 1e-9
    # Snell back into the lens (water -> lens: sin grows by c1/c2); rays beyond the critical angle do not land
    land = np.isfinite(tx_)
    assert land.sum() > 50 and not (land & ~back).any()
    a2 = np.arctan2(ix, iz)
    t2x, t2z = _unit(*lens.tangent(a2))
    s_w = doutx * t2x + doutz * t2z
    d3x, d3z = _unit(tx_ - ix, lens.d - iz)
    s_l = d3x * t2x + d3z * t2z
    assert np.abs(s_l - (lens.c1 / lens.c2) * s_w)[land].max() < 1e-9
    assert (np.abs((lens.c1 / lens.c2) * s_w)[back & ~land] > 1.0 - 1e-9).all()    # total internal reflection
