"""Inputs of the reference scripts, rebuilt with the same NumPy expressions.

SURVEY.md H4: `alpha`, `x_a` must be bit-identical to what main_rt.py builds
(e.g. alpha[452] is -1.11e-16, not 0), so they are produced on the host by the
very expressions of main_rt.py:449-479 instead of being re-derived.
"""
import numpy as np

# main_rt.py:449-460
C1 = np.float64(6400)
C2 = np.float64(1483)
L0 = np.float64(0.12156646438729327)
H0 = np.float64(0.08843353561270673)
D = L0 + H0
ALPHA_MAX = np.float64(50.62033040986099 * (np.pi / 180))
NUM_ELEMENTS = np.int64(64)
PITCH = np.float64(0.0006)
NUM_ALPHA_POINTS_RT = np.int64(181 * 5)        # main_rt.py:462
NUM_ALPHA_POINTS_COMPARE = np.int64(181 * 10)  # main_compare.py:463


def transducer(num_elements=NUM_ELEMENTS, pitch=PITCH, d=D, insert_center=True):
    """main_rt.py:469-474: centred linear array plus the virtual centre element
    inserted at index num_elements // 2 (32 in the reference)."""
    x_a = np.arange(num_elements, dtype=np.float64) * pitch
    x_a = x_a - np.mean(x_a)
    if insert_center:
        x_aux = list(x_a)
        x_aux.insert(int(num_elements) // 2, np.float64(0.0))
        x_a = np.asarray(x_aux, dtype=np.float64)
    z_a = np.ones_like(x_a) * d
    return x_a, z_a


def launch_angles(num_alpha_points=NUM_ALPHA_POINTS_RT, alpha_max=ALPHA_MAX):
    """main_rt.py:479."""
    return np.linspace(-alpha_max, alpha_max, num_alpha_points)


def focus_line(num_alpha_points=NUM_ALPHA_POINTS_RT, d=D):
    """main_rt.py:477: zf, the array line every ray is landed on."""
    return np.ones((int(num_alpha_points),), dtype=np.float64) * d


def sweep_grid():
    """main_rt.py:464-467: the 10 x 21 (r_outer, pipe_offset) geometry sweep in
    the reference's row order (radius-major)."""
    radii, offsets = [], []
    for r in range(1, 10 + 1):
        for p_o in range(-10, 10 + 1):
            radii.append(np.float64(r * 1e-2))
            offsets.append(np.float64(p_o * 1e-3))
    return np.asarray(radii), np.asarray(offsets)


def synthetic_grid(n_radius, n_offset, r_range=(0.01, 0.10), off_range=(-0.01, 0.01)):
    """deterministic synthetic geometry grids over the
    reference's own parameter ranges, radius-major like the reference sweep."""
    r = np.linspace(r_range[0], r_range[1], int(n_radius))
    o = np.linspace(off_range[0], off_range[1], int(n_offset))
    return np.repeat(r, len(o)), np.tile(o, len(r))
