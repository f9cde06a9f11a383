#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

It imports /root/reference/main_rt.py UNMODIFIED (SURVEY.md appendix C recipe:
stub matplotlib, set the module globals the functions read), runs
`shoot_rays` (main_rt.py:337-441) and replicates the `__main__` table assembly
(main_rt.py:487-501) and the per-ray record of main_compare.py:508-524.
It also parses the reference's own fixtures (database_2.csv, compare.csv) into
compact arrays.  Nothing from the reference's sources is copied: only numbers.

Outputs (all np.savez_compressed):
  database_2.npz   parsed /root/reference/database_2.csv (13 650 rows)
  compare.npz      parsed /root/reference/compare.csv (1 810 rows)
  sweep.npz        live reference over the 210-geometry main_rt.py sweep:
                   table (hit, k_hit, tof) + per-ray idx / margin / NaN masks
  rays.npz         full per-ray arrays of `shoot_rays` for a geometry subset
  compare_live.npz live reference for the main_compare.py case (K=N=1810)
  shooters.npz     off-centre shooters (parity "unpinned by fixture", live only)
  array128.npz     live reference with a 128-element array (E = 129: BASELINE configs 3/4) over 96 geometries
                   of the 1024 x 1024 synthetic grid (`python make_golden.py array128` regenerates only this one)
"""
import csv
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = lambda name: (lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    spec = importlib.util.spec_from_file_location("ref_main_rt", os.path.join(REF, "main_rt.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref.c1, ref.c2, ref.c3 = np.float64(6400), np.float64(1483), np.float64(5600)
    ref.l0, ref.h0 = np.float64(0.12156646438729327), np.float64(0.08843353561270673)
    ref.d = ref.l0 + ref.h0
    ref.alpha_max = np.float64(50.62033040986099 * (np.pi / 180))
    ref.num_elements, ref.pitch = np.int64(64), np.float64(0.0006)
    return ref


def array_and_alpha(ref, n_alpha):
    """main_rt.py:469-479 evaluated with the same NumPy expressions."""
    ref.num_alpha_points = np.int64(n_alpha)
    x_a = np.arange(ref.num_elements, dtype=np.float64) * ref.pitch
    x_a = x_a - np.mean(x_a)
    x_aux = list(x_a)
    x_aux.insert(int(ref.num_elements) // 2, np.float64(0.0))   # 32 in the reference (main_rt.py:472)
    x_a = np.asarray(x_aux, dtype=np.float64)
    z_a = np.ones_like(x_a) * ref.d
    zf = np.ones((ref.num_alpha_points,), dtype=np.float64) * ref.d
    alpha = np.linspace(-ref.alpha_max, ref.alpha_max, ref.num_alpha_points)
    return x_a, z_a, zf, alpha


def reference_idx(ref, x_a0, z_a0, alpha, res):
    """First-sign-change index per ray exactly as the reference finds it
    (polyfit line, main_rt.py:385-390 -> :73-99), plus the noise margin
    min(|diffs[idx]|, |diffs[idx+1]|) and the sign-change count.
    idx = -1: no sign change (lines :84-96); -2: line is NaN."""
    n = len(alpha)
    x_p, z_p = res["lens_1_x"], res["lens_1_z"]
    x_q, z_q = res["pipe_x"], res["pipe_z"]
    # a_l, b_l are not returned by shoot_rays: rebuild them through the
    # reference's own helpers (main_rt.py:341-376).
    phi_ap = np.arctan2(z_a0 - z_p, x_a0 - x_p)
    phi_pq, _ = ref.refraction(phi_ap, ref.dz_dx_from_alpha(alpha), ref.c1, ref.c2)
    phi_l, _ = ref.reflection(phi_pq, ref.dzdx_pipe(x_q, ref.r_outer))
    a_l = np.tan(phi_l)
    b_l = z_q - a_l * x_q
    idx = np.full(n, -2, dtype=np.int32)
    margin = np.full(n, np.nan)
    nchg = np.zeros(n, dtype=np.int32)
    for ray in range(n):
        if not (np.isfinite(x_q[ray]) and np.isfinite(a_l[ray]) and np.isfinite(b_l[ray])):
            continue
        xx = np.linspace(x_q[ray] - 0.15, x_q[ray] + 0.15, n)
        zz = a_l[ray] * xx + b_l[ray]
        m, b = np.polyfit(xx, zz, 1)
        diffs = z_p - (m * x_p + b)
        sc = np.where(np.diff(np.sign(diffs)) != 0)[0]
        nchg[ray] = len(sc)
        if len(sc) == 0:
            idx[ray] = -1
            continue
        idx[ray] = sc[0]
        margin[ray] = min(abs(diffs[sc[0]]), abs(diffs[sc[0] + 1]))
    return idx, margin, nchg, a_l, b_l


def assemble_table(ref, x_a, z_a, element_idx, res, atol):
    """main_rt.py:487-501 (first matching ray wins; legs summed left to right)."""
    e = len(x_a)
    hit = np.zeros(e, dtype=bool)
    k_hit = np.full(e, -1, dtype=np.int32)
    tof = np.zeros(e, dtype=np.float64)
    tx = res["target_x"]
    for elem_idx, elem_x in enumerate(x_a):
        close = np.isclose(tx, elem_x, atol=atol)
        ks = np.nonzero(close)[0]
        if len(ks) == 0:
            continue
        ray = int(ks[0])
        t = ref.dist(x_a[element_idx], z_a[element_idx], res["lens_1_x"][ray], res["lens_1_z"][ray]) / ref.c1
        t += ref.dist(res["lens_1_x"][ray], res["lens_1_z"][ray], res["pipe_x"][ray], res["pipe_z"][ray]) / ref.c2
        t += ref.dist(res["pipe_x"][ray], res["pipe_z"][ray], res["lens_2_x"][ray], res["lens_2_z"][ray]) / ref.c2
        t += ref.dist(res["lens_2_x"][ray], res["lens_2_z"][ray], res["target_x"][ray], res["target_z"][ray]) / ref.c1
        hit[elem_idx], k_hit[elem_idx], tof[elem_idx] = True, ray, t
    return hit, k_hit, tof


def legs(ref, x0, z0, res):
    """main_compare.py:514-517, vectorised over rays (dist is elementwise)."""
    t1 = ref.dist(x0, z0, res["lens_1_x"], res["lens_1_z"]) / ref.c1
    t2 = ref.dist(res["lens_1_x"], res["lens_1_z"], res["pipe_x"], res["pipe_z"]) / ref.c2
    t3 = ref.dist(res["pipe_x"], res["pipe_z"], res["lens_2_x"], res["lens_2_z"]) / ref.c2
    t4 = ref.dist(res["lens_2_x"], res["lens_2_z"], res["target_x"], res["target_z"]) / ref.c1
    return np.stack([t1, t2, t3, t4])


def parse_fixture_csvs():
    with open(os.path.join(REF, "database_2.csv"), newline="") as f:
        rows = list(csv.reader(f))
    assert rows[0] == ["elem_idx", "offset", "radius", "hitted", "tof_total"]
    body = rows[1:]
    np.savez_compressed(
        os.path.join(OUT, "database_2.npz"),
        elem_idx=np.array([int(r[0]) for r in body], dtype=np.int32),
        offset=np.array([float(r[1]) for r in body]),
        radius=np.array([float(r[2]) for r in body]),
        hitted=np.array([r[3] == "True" for r in body]),
        tof_total=np.array([float(r[4]) for r in body]),
        # first rows verbatim (strings), to pin the CSV text format
        head_text=np.array([",".join(r) for r in rows[:4]] + [",".join(r) for r in rows[32:36]]),
    )
    with open(os.path.join(REF, "compare.csv"), newline="") as f:
        rows = list(csv.reader(f))
    assert rows[0] == ["alpha", "offset", "radius", "hitted", "tof_1", "tof_2", "tof_3", "tof_4"]
    body = rows[1:]
    np.savez_compressed(
        os.path.join(OUT, "compare.npz"),
        alpha=np.array([float(r[0]) for r in body]),
        offset=np.array([float(r[1]) for r in body]),
        radius=np.array([float(r[2]) for r in body]),
        hitted=np.array([r[3] == "True" for r in body]),
        tof=np.array([[float(r[4 + i]) for r in body] for i in range(4)]),
        head_text=np.array([",".join(r) for r in rows[:3]] + [",".join(r) for r in rows[904:908]]),
    )


RAY_KEYS = ["pipe_x", "pipe_z", "lens_2_x", "lens_2_z", "target_x"]


def array128(ref):
    """BASELINE configs 3/4: the same script with num_elements = 128 (E = 129, shooter = inserted centre element 64)
    on 96 geometries of the 1024 x 1024 synthetic grid (radius-major np.linspace grid of SURVEY 8(d), flat indices
    spread evenly with a stride coprime to 1024 so radii and offsets both vary)."""
    ref.num_elements = np.int64(128)
    x_a, z_a, zf, alpha = array_and_alpha(ref, 181 * 5)
    assert len(x_a) == 129 and x_a[64] == 0.0
    rr = np.linspace(0.01, 0.10, 1024)
    oo = np.linspace(-0.01, 0.01, 1024)
    flat = (np.arange(96, dtype=np.int64) * 10939 + 517) % (1024 * 1024)
    flat = np.sort(flat)
    out = {k: [] for k in ("hit", "k_hit", "tof", "idx", "margin", "nan_target")}
    for g in flat:
        ref.r_outer, ref.pipe_offset = np.float64(rr[g // 1024]), np.float64(oo[g % 1024])
        res = ref.shoot_rays(x_a[64], z_a[64], zf, alpha, plot=False)
        idx, margin, nchg, _, _ = reference_idx(ref, x_a[64], z_a[64], alpha, res)
        hit, k_hit, tof = assemble_table(ref, x_a, z_a, 64, res, atol=1e-6)
        out["hit"].append(hit); out["k_hit"].append(k_hit); out["tof"].append(tof)
        out["idx"].append(idx.astype(np.int16)); out["margin"].append(margin.astype(np.float32))
        out["nan_target"].append(np.isnan(res["target_x"]))
        print(f"array128 g={g} hits={hit.sum()}", flush=True)
    np.savez_compressed(os.path.join(OUT, "array128.npz"), x_a=x_a, alpha=alpha, flat_index=flat,
                        radius=rr[flat // 1024], offset=oo[flat % 1024],
                        nan_target=np.packbits(np.array(out.pop("nan_target")), axis=1),
                        **{k: np.array(v) for k, v in out.items()})
    ref.num_elements = np.int64(64)


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    if sys.argv[1:] == ["array128"]:
        array128(ref)
        return
    parse_fixture_csvs()

    # ---- the main_rt.py sweep (main_rt.py:464-501), K = N = 905 -------------
    x_a, z_a, zf, alpha = array_and_alpha(ref, 181 * 5)
    radii, offsets = [], []
    tab_hit, tab_k, tab_tof = [], [], []
    all_idx, all_margin, all_nchg, nan_tx, nan_l2 = [], [], [], [], []
    keep = {}  # geometry subset with full per-ray arrays
    subset = {(1, -10), (1, 0), (1, 10), (2, 3), (3, -7), (4, 0), (4, 4), (5, -1), (6, 9),
              (7, -10), (8, 2), (9, -4), (10, 0), (10, 10), (10, -10), (3, 1)}
    lens = None
    for r in range(1, 10 + 1):
        for p_o in range(-10, 10 + 1):
            ref.r_outer = np.float64(r * 1e-2)
            ref.pipe_offset = np.float64(p_o * 1e-3)
            res = ref.shoot_rays(x_a[32], z_a[32], zf, alpha, plot=False)
            idx, margin, nchg, a_l, b_l = reference_idx(ref, x_a[32], z_a[32], alpha, res)
            hit, k_hit, tof = assemble_table(ref, x_a, z_a, 32, res, atol=1e-6)
            radii.append(ref.r_outer); offsets.append(ref.pipe_offset)
            tab_hit.append(hit); tab_k.append(k_hit); tab_tof.append(tof)
            all_idx.append(idx); all_margin.append(margin); all_nchg.append(nchg)
            nan_tx.append(np.isnan(res["target_x"])); nan_l2.append(np.isnan(res["lens_2_x"]))
            if lens is None:
                lens = (res["lens_1_x"].copy(), res["lens_1_z"].copy())
            if (r, p_o) in subset:
                keep[(r, p_o)] = {k: res[k].copy() for k in RAY_KEYS}
                keep[(r, p_o)]["a_l"], keep[(r, p_o)]["b_l"] = a_l, b_l
            print(f"sweep r={r} p_o={p_o} hits={hit.sum()}", flush=True)
    np.savez_compressed(
        os.path.join(OUT, "sweep.npz"),
        x_a=x_a, alpha=alpha, zf0=zf[0], radius=np.array(radii), offset=np.array(offsets),
        hit=np.array(tab_hit), k_hit=np.array(tab_k), tof=np.array(tab_tof),
        idx=np.array(all_idx).astype(np.int16), margin=np.array(all_margin).astype(np.float32),
        nchg=np.array(all_nchg).astype(np.int8),
        nan_target=np.packbits(np.array(nan_tx), axis=1), nan_lens2=np.packbits(np.array(nan_l2), axis=1),
        lens_x=lens[0], lens_z=lens[1],
    )
    gl = sorted(keep)
    np.savez_compressed(
        os.path.join(OUT, "rays.npz"),
        r_index=np.array([g[0] for g in gl]), p_index=np.array([g[1] for g in gl]),
        radius=np.array([np.float64(g[0] * 1e-2) for g in gl]),
        offset=np.array([np.float64(g[1] * 1e-3) for g in gl]),
        **{k: np.array([keep[g][k] for g in gl]) for k in RAY_KEYS + ["a_l", "b_l"]},
    )

    # ---- the main_compare.py case (main_compare.py:463-481, :508-521) -------
    x_a2, z_a2, zf2, alpha2 = array_and_alpha(ref, 181 * 10)
    ref.r_outer, ref.pipe_offset = np.float64(0.037), np.float64(0.0038)
    res = ref.shoot_rays(x_a2[32], z_a2[32], zf2, alpha2, plot=False)
    idx, margin, nchg, a_l, b_l = reference_idx(ref, x_a2[32], z_a2[32], alpha2, res)
    hitted = np.zeros(len(alpha2), dtype=bool)
    for ray in range(len(alpha2)):
        for elem_x in x_a2:
            if np.isclose(res["target_x"][ray], elem_x, atol=1e-4):
                hitted[ray] = True
                break
    np.savez_compressed(
        os.path.join(OUT, "compare_live.npz"),
        alpha=alpha2, x_a=x_a2, hitted=hitted, legs=legs(ref, x_a2[32], z_a2[32], res),
        idx=idx, margin=margin.astype(np.float32), nchg=nchg.astype(np.int8), a_l=a_l, b_l=b_l,
        lens_x=res["lens_1_x"], lens_z=res["lens_1_z"],
        **{k: res[k] for k in RAY_KEYS},
    )

    # ---- off-centre shooters (no fixture in the reference; live only) -------
    x_a, z_a, zf, alpha = array_and_alpha(ref, 181 * 5)
    out = {"shooter": [], "radius": [], "offset": [], "hit": [], "k_hit": [], "tof": [], "idx": [], "margin": []}
    for k in RAY_KEYS:
        out[k] = []
    for (s, rr, oo) in [(0, 0.08, 0.002), (20, 0.06, -0.003), (45, 0.09, 0.0), (64, 0.1, 0.007), (33, 0.05, 0.001)]:
        ref.r_outer, ref.pipe_offset = np.float64(rr), np.float64(oo)
        try:
            res = ref.shoot_rays(x_a[s], z_a[s], zf, alpha, plot=False)
        except np.linalg.LinAlgError:
            print(f"shooter {s} r={rr} off={oo}: reference crashed (ray misses the pipe) - skipped")
            continue
        idx, margin, nchg, _, _ = reference_idx(ref, x_a[s], z_a[s], alpha, res)
        hit, k_hit, tof = assemble_table(ref, x_a, z_a, s, res, atol=1e-6)
        out["shooter"].append(s); out["radius"].append(rr); out["offset"].append(oo)
        out["hit"].append(hit); out["k_hit"].append(k_hit); out["tof"].append(tof)
        out["idx"].append(idx); out["margin"].append(margin.astype(np.float32))
        for k in RAY_KEYS:
            out[k].append(res[k])
        print(f"shooter {s} r={rr} off={oo}: hits={hit.sum()}")
    np.savez_compressed(os.path.join(OUT, "shooters.npz"), x_a=x_a, alpha=alpha,
                        **{k: np.array(v) for k, v in out.items()})
    array128(ref)


if __name__ == "__main__":
    main()
