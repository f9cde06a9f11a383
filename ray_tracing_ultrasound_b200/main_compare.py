"""Drop-in for the reference's main_compare.py on the B200 engine.

main_compare.py:449-553: trace one geometry (r_outer = 0.037, pipe_offset =
0.0038) with K = N = 1810, write `compare.csv` (alpha, offset, radius, hitted,
tof_1..tof_4; hit tolerance atol = 1e-4), read the matching slice of
`database.csv`, print the per-ray error vector, `num_hitted` and the MSE.

`database.csv` was stripped from the reference checkout (.MISSING_LARGE_BLOBS);
`make_database()` regenerates a file with the schema main_compare.py:483-500
expects (1 810 rows per (offset, radius) pair) over any geometry list, using the
engine's per-ray records.

    python -m ray_tracing_ultrasound_b200.main_compare            # needs database.csv
    python -m ray_tracing_ultrasound_b200.main_compare --make-database
"""
import csv
import sys

import numpy as np

from . import reference_setup as rs
from .engine import TofEngine

CSV_HEADER = ["alpha", "offset", "radius", "hitted", "tof_1", "tof_2", "tof_3", "tof_4"]

# main_compare.py:450-466
c1, c2 = rs.C1, rs.C2
l0, h0 = rs.L0, rs.H0
d = l0 + h0
alpha_max = rs.ALPHA_MAX
num_elements, pitch = rs.NUM_ELEMENTS, rs.PITCH
num_alpha_points = rs.NUM_ALPHA_POINTS_COMPARE
r_outer = np.float64(0.037)
pipe_offset = np.float64(0.0038)


def trace(radii, offsets, element_idx=32, engine=None):
    """Per-ray records (four legs, landing point) of the K = 1810 shot for each geometry."""
    x_a, z_a = rs.transducer(num_elements, pitch, d)
    zf = rs.focus_line(num_alpha_points, d)
    alpha = np.linspace(-alpha_max, alpha_max, num_alpha_points)
    eng = engine or TofEngine(0)
    eng.setup(alpha, zf, x_a[element_idx], z_a[element_idx], x_a, c1=c1, c2=c2, l0=l0, h0=h0, d=d, atol=1e-4)
    t = eng.tof_table(radii, offsets, records=("legs", "target_x"))
    # main_compare.py:518-521: a ray is "hitted" if it lands on ANY element (np.isclose, atol 1e-4)
    tx = t.records["target_x"]
    hitted = np.isclose(tx[:, :, None], x_a[None, None, :], atol=1e-4).any(axis=2)
    return alpha, t.records["legs"], hitted


def write_rays(filename, alpha, offset, radius, legs, hitted, header=True, mode="w"):
    """main_compare.py:504-524 row format (CRLF, `nan` for NaN legs)."""
    with open(filename, mode, newline="") as csvfile:
        csvwriter = csv.writer(csvfile)
        if header:
            csvwriter.writerow(CSV_HEADER)
        for ray in range(len(alpha)):
            csvwriter.writerow([alpha[ray], offset, radius, bool(hitted[ray]), legs[0, ray], legs[1, ray],
                                legs[2, ray], legs[3, ray]])


def make_database(filename="database.csv", radii=None, offsets=None, engine=None):
    """Regenerate database.csv (schema main_compare.py:483-500) over a geometry list; by
    default the 10 x 21 grid of main_rt.py:464-467 plus the geometry main_compare.py uses."""
    if radii is None:
        radii, offsets = rs.sweep_grid()
        radii, offsets = np.append(radii, r_outer), np.append(offsets, pipe_offset)
    alpha, legs, hitted = trace(radii, offsets, engine=engine)
    for g in range(len(radii)):
        write_rays(filename, alpha, np.float64(offsets[g]), np.float64(radii[g]), legs[g], hitted[g],
                   header=(g == 0), mode="w" if g == 0 else "a+")
    return len(radii) * len(alpha)


def process_tofs(legs):
    """main_compare.py:526-534: row sum of tof_1..4, NaN if any leg is NaN (pandas sum order)."""
    total = ((legs[0] + legs[1]) + legs[2]) + legs[3]
    return np.where(np.isnan(legs).any(axis=0), np.nan, total)


def main(database="database.csv", filename_compare="compare.csv", engine=None):
    import pandas as pd
    alpha, legs, hitted = trace([r_outer], [pipe_offset], engine=engine)
    df = pd.read_csv(database)
    # main_compare.py:489-500: snap to the closest offset / radius present in the database
    off_col, rad_col = df["offset"].to_numpy(), df["radius"].to_numpy()
    offset_closest_value = off_col[np.argmin(np.abs(off_col - pipe_offset))]
    radius_closest_value = rad_col[np.argmin(np.abs(rad_col - r_outer))]
    filtered_df = df[(df["offset"] == offset_closest_value) & (df["radius"] == radius_closest_value)]
    write_rays(filename_compare, alpha, pipe_offset, r_outer, legs[0], hitted[0])
    tofs_database = process_tofs(filtered_df[["tof_1", "tof_2", "tof_3", "tof_4"]].to_numpy(dtype=np.float64).T)
    df_compare = pd.read_csv(filename_compare)
    num_hitted = df_compare["hitted"].sum()
    tofs_compare = process_tofs(df_compare[["tof_1", "tof_2", "tof_3", "tof_4"]].to_numpy(dtype=np.float64).T)
    individual_errors = tofs_database - tofs_compare
    print(individual_errors)
    print(f"{num_hitted = }")
    non_nan_mask = ~np.isnan(individual_errors)
    mse = np.sum(np.square(individual_errors[non_nan_mask])) / num_hitted
    print(f"{mse = }")
    return individual_errors, num_hitted, mse


if __name__ == "__main__":
    if "--make-database" in sys.argv:
        print(make_database(), "rows written to database.csv")
    else:
        main()
