"""Drop-in for the reference's main_rt.py on the B200 engine.

Same module-level "globals" configuration, same `shoot_rays(x_a, z_a, z_f,
alpha, plot=True)` signature and dict keys (main_rt.py:337, :432-441), and the
same `database_2.csv` rows (main_rt.py:484-504) -- but the 210-geometry sweep
is ONE batched call into the CUDA table kernel instead of 210 x 905 Python
iterations.

    python -m ray_tracing_ultrasound_b200.main_rt [database_2.csv]

Differences kept on purpose: `plot` defaults to True as in the reference but
matplotlib is imported lazily and only then; a ray that misses the pipe gives
NaN instead of the reference's LinAlgError crash.
"""
import csv
import sys
import warnings

import numpy as np

from . import reference_setup as rs
from .engine import TofEngine

# ---- the module globals of main_rt.py:449-467 (set them like the reference does) ----
c1 = rs.C1
c2 = rs.C2
c3 = np.float64(5600)  # defined and unused in the reference as well
l0 = rs.L0
h0 = rs.H0
d = l0 + h0
alpha_max = rs.ALPHA_MAX
num_elements = rs.NUM_ELEMENTS
pitch = rs.PITCH
num_alpha_points = rs.NUM_ALPHA_POINTS_RT
r_outer = np.float64(0.04)
pipe_offset = np.float64(0.0)

_engine = None
_engine_key = None


def _get_engine(x_a, z_a, z_f, alpha, x_rx=None, atol=1e-6):
    """One engine per process; tofrt_setup is redone only when the shot changes."""
    global _engine, _engine_key
    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    z_f = np.ascontiguousarray(z_f, dtype=np.float64)
    x_rx = np.ascontiguousarray([x_a] if x_rx is None else x_rx, dtype=np.float64)
    key = (alpha.tobytes(), z_f.tobytes(), x_rx.tobytes(), float(x_a), float(z_a), float(c1), float(c2),
           float(l0), float(h0), float(d), float(atol))
    if _engine is None:
        _engine = TofEngine(0)
    if key != _engine_key:
        _engine.setup(alpha, z_f, x_a, z_a, x_rx, c1=c1, c2=c2, l0=l0, h0=h0, d=d, atol=atol)
        _engine_key = key
    return _engine


def dist(x1, z1, x2, z2):
    """main_rt.py:444-445."""
    return np.sqrt((x1 - x2) ** 2 + (z1 - z2) ** 2)


def shoot_rays(x_a, z_a, z_f, alpha, plot=True):
    """main_rt.py:337-441 for the current globals (r_outer, pipe_offset, c1, ...)."""
    alpha = np.asarray(alpha, dtype=np.float64)
    z_f = np.asarray(z_f, dtype=np.float64)
    if len(alpha) < 2:
        raise ValueError("Curve needs at least two points.")  # main_rt.py:26-27
    if len(alpha) != len(z_f):
        raise ValueError("Coordinate arrays must have the same length.")  # main_rt.py:28-29
    eng = _get_engine(x_a, z_a, z_f, alpha)
    res = eng.shoot_rays(r_outer, pipe_offset)
    if plot:
        _plot(x_a, z_a, res)
    return res


def _plot(x_a, z_a, res):
    """Figure of main_rt.py:407-430 (every 10th ray); optional, lazy matplotlib."""
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        warnings.warn("matplotlib is not installed: shoot_rays(plot=True) skips the figure")
        return
    plt.figure()
    ang = np.linspace(-np.pi / 2, np.pi / 2, len(res["lens_1_x"]))
    plt.plot(res["lens_1_x"], res["lens_1_z"], color="red", label="Refracting surface")
    plt.plot(r_outer * np.sin(ang) + pipe_offset, r_outer * np.cos(ang), color="blue", label="Pipe")
    for ray in range(0, len(res["lens_1_x"]), 10):
        plt.plot([x_a, res["lens_1_x"][ray]], [z_a, res["lens_1_z"][ray]], "C0")
        plt.plot([res["lens_1_x"][ray], res["pipe_x"][ray]], [res["lens_1_z"][ray], res["pipe_z"][ray]], "C1")
        plt.plot([res["pipe_x"][ray], res["lens_2_x"][ray]], [res["pipe_z"][ray], res["lens_2_z"][ray]], "C2")
        plt.plot([res["lens_2_x"][ray], res["target_x"][ray]], [res["lens_2_z"][ray], res["target_z"][ray]], "C3")
    plt.title(f"Element at {x_a} m shooting")
    plt.xlim([-0.1, 0.1])
    plt.axis("equal")
    plt.legend()
    plt.show()


def sweep_table(element_idx=32):
    """main_rt.py:464-501 as one batched call: returns (x_a, radii, offsets, TofTable)."""
    x_a, z_a = rs.transducer(num_elements, pitch, d)
    zf = rs.focus_line(num_alpha_points, d)
    alpha = np.linspace(-alpha_max, alpha_max, num_alpha_points)
    radii, offsets = rs.sweep_grid()
    eng = _get_engine(x_a[element_idx], z_a[element_idx], zf, alpha, x_rx=x_a, atol=1e-6)
    return x_a, radii, offsets, eng.tof_table(radii, offsets)


def write_rows(filename, x_a, radii, offsets, table):
    """main_rt.py:487-504: append one CRLF row per (geometry, element); `0` when not hit."""
    with open(filename, "a+", newline="") as csvfile:
        csvwriter = csv.writer(csvfile)
        for g in range(len(radii)):
            for elem_idx in range(len(x_a)):
                hit = bool(table.k_hit[g, elem_idx] >= 0)
                csvwriter.writerow([elem_idx, offsets[g], radii[g], hit, table.tof[g, elem_idx] if hit else 0])


def main(filename="database_2.csv"):
    x_a, radii, offsets, table = sweep_table()
    write_rows(filename, x_a, radii, offsets, table)
    return table


if __name__ == "__main__":
    main(*sys.argv[1:2])
