#!/usr/bin/env python
"""profiles/sass_r2_excerpts.txt: cuobjdump -sass of the shipped libtofrt.so -- arch, per-kernel instruction statistics,
the TMA bulk copy + mbarrier instructions of the table kernel, one bracket probe, the literal scan's inner loop.
    python tools/sass_excerpts.py [out.txt]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

lib = os.path.join(ROOT, "ray_tracing_ultrasound_b200", "libtofrt.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
out = [f"# SASS excerpts of ray_tracing_ultrasound_b200/libtofrt.so (cuobjdump -sass), sources with hash {bench.source_hash()}",
       "# arch line, per-kernel static instruction statistics, the TMA bulk copy + mbarrier instructions of the table kernel,",
       "# one bracket probe of the guided search, the inner loop of the literal scan (LDS.128 + DMUL + DADD + DADD + VIMNMX3 per two samples).", ""]
out += ["arch = " + re.search(r"arch = (\S+)", txt).group(1), ""]
ins = lambda p: [l for l in p.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
clean = lambda l: re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip()
out.append("## kernels (static instruction counts; no HMMA / UTC* tensor-core instructions anywhere: not a contraction)")
for p in parts[1:]:
    L = ins(p)
    c = lambda s: sum(s in l for l in L)
    out.append(p.split("\n")[0])
    out.append(f"    instructions {len(L)}  DFMA {c('DFMA')} DMUL {c('DMUL')} DADD {c('DADD')} DSETP {c('DSETP')} MUFU.RCP64H {c('MUFU.RCP64H')} "
               f"MUFU.RSQ64H {c('MUFU.RSQ64H')} FFMA {c('FFMA')} LDS {c(' LDS')} LDG {c('LDG')} UBLKCP {c('UBLKCP')} SYNCS {c('SYNCS')} "
               f"BAR {c('BAR.SYNC')} VIMNMX3 {c('VIMNMX3')} HMMA/UTC* {c('HMMA') + c('UTC')}")


def kernel(sub):
    for p in parts[1:]:
        if sub in p.split("\n")[0]:
            return ins(p)


g = kernel("tofrt_table_kernelILb0ELb1ELb0ELb1ELb0ELb0")
out.append("\n## tofrt_table_kernel<0,1,0,1,0,0> (guided, vector form): TMA bulk copies completing on an mbarrier")
out += [clean(l) for l in g if "UBLKCP" in l or "SYNCS" in l]
out.append("\n## one bracket probe of the guided search (lens sample + chord slope from shared memory, one DFMA, two DSETP)")
for i in [i for i, l in enumerate(g) if "LDS.128" in l]:
    win = g[i:i + 14]
    if any("DFMA" in w for w in win) and sum("DSETP" in w for w in win) >= 2 and any("LDS.64" in w for w in win):
        out += [clean(w) for w in win]
        break
s = kernel("tofrt_table_kernelILb0ELb0ELb0ELb1ELb0ELb0")
out.append("\n## tofrt_table_kernel<0,0,0,1,0,0> (literal scan): inner loop, 40 instructions of the unrolled 32-sample chunk")
first = [i for i, l in enumerate(s) if "VIMNMX3" in l][0]
out += [clean(w) for w in s[max(0, first - 12):max(0, first - 12) + 40]]
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_r2_excerpts.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path, len(out), "lines")
