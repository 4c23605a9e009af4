"""Builds libvipcup.so in-tree with nvcc for sm_100a (one shared library, plain C ABI, see include/vipcup.h).

Every csrc/*.cu is compiled to build/<name>.o in parallel (only the sources that changed), then linked."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvipcup.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into libvipcup.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libvipcup.so cannot be built")
    extra = os.environ.get("VIP_NVCC_EXTRA", "").split()   # e.g. -DVIP_MBAR_DEBUG (barrier watchdog for kernel bring-up)
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, ".flags")
    flags_now = " ".join(NVCC_FLAGS + extra)
    flags_old = open(stamp).read() if os.path.exists(stamp) else None
    hdr_t = max(os.path.getmtime(h) for h in _headers())

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        fresh = (not force and flags_old == flags_now and os.path.exists(obj)
                 and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t))
        if fresh:
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []), "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    with open(stamp, "w") as f:
        f.write(flags_now)
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *[o for o, _ in results],
                          "-o", LIB + ".tmp"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print("".join(log for _, log in results))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
