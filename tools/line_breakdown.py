#!/usr/bin/env python
"""Per-source-line instruction counts of a kernel: joins an `ncu --set full --import-source on` capture with
`nvdisasm -gi` line info of the SAME libtofrt.so build (see tools/stage_breakdown.py).

    python tools/line_breakdown.py gpurun_out/prof.ncu-rep tofrt_table_kernelILb0ELb1ELb0ELb1 [chunks]

Prints, per line of the kernel's source file (outermost call site), executed warp instructions split into
FP64-pipe and other, per 32-ray chunk when `chunks` is given."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stage_breakdown as sb


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    chunks = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    dis, smp = sb.disasm(kernel), sb.samples(rep)
    if len(dis) != len(smp):
        sys.exit(f"instruction count differs: nvdisasm {len(dis)} vs capture {len(smp)}")
    src = open(os.path.join(sb.CSRC, sb.source_file(kernel))).read().splitlines()
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    for (off, op, line), (_, ns, ni) in zip(dis, smp):
        key = (os.path.basename(line[0]), line[1]) if line else ("?", 0)
        a = agg[key]
        a[0] += ns
        a[1 if sb.FP64.match(op) else 2] += ni
        a[3][op.split(".")[0]] += ni
    tot = sum(a[1] + a[2] for a in agg.values())
    print(f"total {tot / chunks:.1f} per chunk")
    for key in sorted(agg):
        a = agg[key]
        n = a[1] + a[2]
        if n / chunks < 0.5:
            continue
        text = src[key[1] - 1].strip()[:90] if key[0] == sb.source_file(kernel) and 0 < key[1] <= len(src) else ""
        top = " ".join(f"{k}:{v / chunks:.0f}" for k, v in a[3].most_common(6))
        print(f"{key[0][:22]:22s}:{key[1]:4d} fp64 {a[1] / chunks:6.1f} other {a[2] / chunks:6.1f} smp {a[0]:6d} | {text}\n{'':34s}{top}")


if __name__ == "__main__":
    main()
