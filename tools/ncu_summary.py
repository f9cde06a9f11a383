#!/usr/bin/env python
"""Selected raw metrics of `ncu --set full` captures -> one JSON (what profiles/ncu_*.json hold).
    python tools/ncu_summary.py out.json name1=rep1.ncu-rep name2=rep2.ncu-rep ...

The FIRST capture also fills the top-level keys bench.py reads (kernel_ms, fp64_pipe_pct, issue_slots_busy_pct,
dram_bytes, warp_instructions) and the file is stamped with `source_hash` -- bench.source_hash() of the tree the summary
is made in, i.e. run this right after the capture, before touching csrc/ -- so that bench.py can refuse to quote a
capture of another build."""
import csv
import io
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"gpu__time_duration\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|launch__(block_size|grid_size|"
                  r"registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_registers)|sm__cycles_elapsed\.avg|"
                  r"sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|"
                  r"smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed\.sum|"
                  r"smsp__thread_inst_executed_per_inst_executed\.ratio|sm__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__inst_executed_pipe_fp64\.sum|sm__inst_executed_pipe_fp64\.sum)$")

out = {}
for arg in sys.argv[2:]:
    name, rep = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")]}
    for h, u, v in zip(hdr, units, vals):
        if KEEP.match(h):
            try:
                d[h] = {"unit": u, "value": float(v.replace(",", ""))}
            except ValueError:
                pass
    out[name] = d
first = next(iter(out.values()))
val = lambda k: first.get(k, {}).get("value")
import bench
top = {"source_hash": bench.source_hash(), "file": os.path.basename(sys.argv[2].split("=", 1)[1]),
       "kernel_ms": val("gpu__time_duration.sum"),
       "fp64_pipe_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
       "issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active") or val("sm__issue_active.avg.pct_of_peak_sustained_active"),
       "warp_instructions": val("smsp__inst_executed.sum"),
       "dram_bytes": None if val("dram__bytes_read.sum") is None else
       val("dram__bytes_read.sum") * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[first["dram__bytes_read.sum"]["unit"]] +
       val("dram__bytes_write.sum") * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[first["dram__bytes_write.sum"]["unit"]]}
if first.get("gpu__time_duration.sum", {}).get("unit") in ("us", "usecond"):
    top["kernel_ms"] = top["kernel_ms"] / 1e3
elif first.get("gpu__time_duration.sum", {}).get("unit") in ("ns", "nsecond"):
    top["kernel_ms"] = top["kernel_ms"] / 1e6
elif first.get("gpu__time_duration.sum", {}).get("unit") in ("s", "second"):
    top["kernel_ms"] = top["kernel_ms"] * 1e3
out = dict(top, captures=out)
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print(json.dumps({k: {m: v["value"] for m, v in d.items() if isinstance(v, dict) and ("fp64" in m or "time_duration" in m or "issue_active" in m or "dram__bytes" in m or "warps_active" in m or m == "smsp__inst_executed.sum")} for k, d in out["captures"].items()}, indent=1))
print({k: v for k, v in out.items() if k != "captures"})
