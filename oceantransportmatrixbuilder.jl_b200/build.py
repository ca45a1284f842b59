"""Build recipe for libotmb.so (sm_100a only, in-tree so the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
INCLUDE = HERE.parent / "include"
LIB = HERE / "libotmb.so"
SOURCES = ["ctx.cu", "scan.cu", "geometry.cu", "faceflux.cu", "fused.cu", "fused_v4.cu", "coo.cu", "transport.cu", "redigm.cu", "velocity.cu", "lump.cu", "spmv.cu", "fetch.cu", "comm.cu", "gm.cu", "selftest.cu"]
HEADERS = ["common.cuh", "sphere.cuh", "fused_generic.cuh", "fdiv.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                 # the reference never contracts a*b+c; keep every rounding
    "-Xcompiler", "-fPIC",
    "-shared",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: libotmb.so cannot be built (there is no CPU fallback)")
    return exe


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HEADERS] + [INCLUDE / "otmb.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    """Compile every translation unit for sm_100a (in parallel) and link libotmb.so in-tree."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + [f"-I{INCLUDE}", f"-I{CSRC}"]
    flags += os.environ.get("OTMB_NVCC_EXTRA", "").split()      # e.g. -DOTMB_AB: measurement build (A/B switches, traces)
    if ptxas_info:
        flags += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = objdir / (Path(src).stem + ".o")
        cmd = [nvcc(), *flags, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return obj, r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    log = "".join(out for _, _, out in results)
    if any(rc != 0 for _, rc, _ in results):
        raise RuntimeError("nvcc failed:\n" + log)
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
           *[str(o) for o, _, _ in results], "-ldl", "-o", str(LIB)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout)
    if verbose or ptxas_info:
        print(log + r.stdout, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True, ptxas_info="-v" in sys.argv)
    print(LIB)
