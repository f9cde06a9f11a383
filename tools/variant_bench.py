"""Tuning helper: build launch-shape variants of libtofrt.so (here, on CPU) and time them (on the GPU box).
    python tools/variant_bench.py build      # in the build container
    python tools/variant_bench.py run        # under gpurun
"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {"base": (), "t672x2": ("TOFRT_THREADS_GUIDED=672", "TOFRT_MIN_BLOCKS=2"), "t640x2": ("TOFRT_THREADS_GUIDED=640", "TOFRT_MIN_BLOCKS=2"),
            "t768x2": ("TOFRT_THREADS_GUIDED=768", "TOFRT_MIN_BLOCKS=2"), "t512x3": ("TOFRT_THREADS_GUIDED=512", "TOFRT_MIN_BLOCKS=3")}
VDIR = os.path.join(ROOT, "tools", "_variants")

if sys.argv[1] == "build":
    from ray_tracing_ultrasound_b200 import _build
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        _build.build(force=True, defines=defs, out=os.path.join(VDIR, f"libtofrt_{name}.so"))
        print("built", name)
elif sys.argv[1] == "run":
    for name in VARIANTS:
        lib = os.path.join(VDIR, f"libtofrt_{name}.so")
        if os.path.exists(lib):
            env = dict(os.environ, TOFRT_LIB=lib)
            subprocess.run([sys.executable, __file__, "one", name], env=env)
else:
    import numpy as np, torch
    from ray_tracing_ultrasound_b200 import TofEngine, reference_setup as rs
    x_a, z_a = rs.transducer(num_elements=128); alpha, zf = rs.launch_angles(), rs.focus_line()
    eng = TofEngine(0).setup(alpha, zf, x_a[64], z_a[64], x_a)
    radii, offsets = rs.synthetic_grid(1024, 1024)
    r = torch.from_numpy(radii).cuda(); o = torch.from_numpy(offsets).cuda()
    out = {"variant": sys.argv[2], "info": eng.kernel_info()}
    for mode in ("guided+vector",):
        eng.set_search(mode.split("+")[0])
        eng.set_math("vector" if "+vector" in mode else "angle")
        for _ in range(4): eng.tof_table_cuda(r, o)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); t = eng.tof_table_cuda(r, o); e1.record(); torch.cuda.synchronize()
            ts.append(round(e0.elapsed_time(e1), 2))
        out[mode + "_ms"] = ts
        out[mode + "_sum"] = float(t.tof.sum())
    print(json.dumps(out))
