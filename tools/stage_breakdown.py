#!/usr/bin/env python
"""Where a kernel spends its time, by source stage: joins the per-instruction warp-stall samples of an
`ncu --set full --import-source on` capture with the line info of the SAME libtofrt.so build
(`nvdisasm -gi`: every SASS instruction carries its source line and, for inlined device functions, the
chain of call sites; the outermost call site inside the kernel decides the stage).

    python tools/stage_breakdown.py gpurun_out/prof.ncu-rep tofrt_table_kernelILb0ELb1ELb0 [out.json]

Stages are cut at marker comments of the kernel's source file (csrc/tofrt_table_kernel.cuh,
csrc/tofrt_fermat_kernels.cuh; MARKERS below)."""
import csv
import io
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ray_tracing_ultrasound_b200", "csrc")
SRC_OF = {"tofrt_table_kernel": "tofrt_table_kernel.cuh", "tofrt_fermat3d_kernel": "tofrt_fermat_kernels.cuh"}
LIB = os.path.join(ROOT, "ray_tracing_ultrasound_b200", "libtofrt.so")

# (marker text that starts a stage, stage name); searched in order inside the kernel body
MARKERS = {
    "tofrt_table_kernel": [
        ("tofrt_table_kernel(DevSetup S, TableArgs T)", "kernel prologue (staging, TMA)"),
        ("__shared__ unsigned long long s_base;", "per-geometry setup, chunk loop head, ray-table loads"),
        ("// ---- main_rt.py:351-364 pipe intersection, upper root; :367-376 reflection", "stage 2: pipe intersection + reflection, vector form (main_rt.py:351-376)"),
        ("// ---- main_rt.py:351-364 pipe intersection, upper root ----", "stage 2 in the reference's angle form (main_rt.py:351-376)"),
        ("if (T.inject_m)", "test hook"),
        ("// ---- main_rt.py:78-99 first sign change", "return-path search (main_rt.py:78-99)"),
        ("int flagged = -1;", "exact scans / literal scan + on-sample rule"),
        ("// chord intersection for every lane", "chord intersection (main_rt.py:106-168)"),
        ("// ---- main_rt.py:396-405 refraction back", "stage 4: second refraction, landing point (main_rt.py:396-405)"),
        ("// ---- optional per-ray records", "ray records"),
        ("// ---- main_rt.py:487-501 which receivers", "table assembly (main_rt.py:487-501)"),
        ("// ---- one geometry done", "row store / hit emission"),
        ("if (STATS && T.stats) {", "epilogue"),
    ],
    "tofrt_fermat3d_kernel": [
        ("tofrt_fermat3d_kernel(tofrt_surf3d S, F3dArgs A)", "prologue / work fetch / staging"),
        ("// ---- discrete search over the sampled surface", "discrete search over the surface samples (FP32)"),
        ("// ---- continuous refinement ----\n        const int ia", "projected Newton refinement (FP64)"),
        ("const size_t o = (size_t)e * A.F + f;\n        A.tof[o] = t;", "stores"),
    ],
}
FP64 = re.compile(r"^(D[A-Z]+|MUFU\.(RCP|RSQ)64H)")


def source_file(kernel):
    return SRC_OF[next(k for k in MARKERS if k in kernel)]


def stage_table(kernel):
    text = open(os.path.join(CSRC, source_file(kernel))).read()
    key = next(k for k in MARKERS if k in kernel)
    start = text.index(MARKERS[key][0][0])
    cuts, pos = [], start
    for marker, name in MARKERS[key]:
        pos = text.index(marker, pos)
        cuts.append((text.count("\n", 0, pos) + 1, name))
    return cuts


def disasm(kernel):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        out = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    rows, active, line, chain = [], False, None, []
    for ln in out.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            chain.append((m.group(1), int(m.group(2)), m.group(3) is not None))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            if chain:
                # the last entry that is not itself "inlined at" something is the outermost call site
                outer = [c for c in chain if not c[2]]
                line = (outer[-1] if outer else chain[-1])[:2]
                chain = []
            rows.append((int(m.group(1), 16), m.group(2), line))
    return rows


def samples(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    first = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rd = csv.DictReader(io.StringIO("\n".join(lines[first:])))
    return [(int(r["Address"], 16), int(r["# Samples"]), int(r["Instructions Executed"])) for r in rd]


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    dis, smp = disasm(kernel), samples(rep)
    if len(dis) != len(smp):
        sys.exit(f"instruction count differs: nvdisasm {len(dis)} vs capture {len(smp)} -- the capture is of another build")
    cuts = stage_table(kernel)
    src_name = source_file(kernel)
    agg, tot_s, tot_i = {name: [0, 0, 0] for _, name in cuts}, 0, 0
    unknown = [0, 0]
    for (off, op, line), (_, ns, ni) in zip(dis, smp):
        name = None
        if line and os.path.basename(line[0]) == src_name:
            for ln, nm in cuts:
                if line[1] >= ln:
                    name = nm
        if name is None and line and os.path.basename(line[0]).startswith("tofrt_"):
            name = "out-of-line helpers, division / square-root slow paths"
            agg.setdefault(name, [0, 0, 0])
        if name is None:
            unknown[0] += ns; unknown[1] += ni
            continue
        a = agg[name]
        a[0] += ns; a[1] += ni
        if FP64.match(op):
            a[2] += ni
        tot_s += ns; tot_i += ni
    res = {name: {"time_pct": round(100.0 * a[0] / max(tot_s, 1), 1), "instr_pct": round(100.0 * a[1] / max(tot_i, 1), 1),
                  "fp64_share_pct": round(100.0 * a[2] / max(a[1], 1))} for name, a in agg.items()}
    res["_total"] = {"samples": tot_s, "warp_instructions": tot_i, "unattributed_samples": unknown[0],
                     "unattributed_instructions": unknown[1], "kernel": kernel, "capture": os.path.basename(rep)}
    js = json.dumps(res, indent=1)
    print(js)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(js + "\n")


if __name__ == "__main__":
    main()
