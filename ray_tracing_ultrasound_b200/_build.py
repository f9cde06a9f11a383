"""Builds libtofrt.so (the C-ABI shared library) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TOFRT_LIB") or os.path.join(HERE, "libtofrt.so")  # TOFRT_LIB: tuning variants
SOURCES = [os.path.join(HERE, "csrc", "tofrt_kernels.cu")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in sorted(os.listdir(os.path.join(HERE, "csrc")))
           if f.endswith((".h", ".cuh"))] + [os.path.join(os.path.dirname(HERE), "include", "tofrt.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # NumPy never fuses mul+add; fused ops are explicit fma() only
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-shared",
]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libtofrt.so cannot be built")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu -> libtofrt.so.  Cross-compiles without a GPU.  Safe to call from several
    processes at once (torchrun ranks): an exclusive file lock serialises the compile."""
    import fcntl
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not is_stale():
        return LIB_PATH
    with open(out + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and out == LIB_PATH and not is_stale():
                return LIB_PATH  # another process built it while we waited
            tmp = out + f".tmp{os.getpid()}"
            cmd = ([nvcc_path()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
                   + ["-o", tmp] + SOURCES)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, out)  # atomic: a reader never sees a half-written library
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return out
