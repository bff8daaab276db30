"""Builds regcm_b200/csrc -> regcm_b200/libmoloch_b200.so (nvcc, sm_100a only).

The library is built IN-TREE so that it travels to the GPU box with the
repository snapshot; nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmoloch_b200.so")
# the same sources with FMA contraction allowed ("fast" mode: not bit-identical to the reference's arithmetic,
# within the tolerances of SURVEY.md 8(c); tests/test_gpu_zz_fastmode.py; bench.py --mode fast)
LIB_FAST = os.path.join(HERE, "libmoloch_b200_fast.so")
SOURCES = ["kernels.cu", "kernels_sound.cu", "kernels_waf.cu", "kernels_bdy.cu", "halo.cu", "capi.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "geo.h"), os.path.join(CSRC, "bdy_cells.h"),
           os.path.join(HERE, "..", "include", "moloch_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-for-bit parity with a non-contracting FP64 evaluation of the Fortran
    "-fmad=false",
    # ... and of the host-evaluated scalars that feed the kernels (boundary time weights, spectral weights)
    "-Xcompiler", "-ffp-contract=off",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _nccl_include() -> list[str]:
    for d in ("/usr/include",):
        if os.path.exists(os.path.join(d, "nccl.h")):
            return []
    try:
        import nvidia.nccl  # type: ignore
        return ["-I", os.path.join(os.path.dirname(nvidia.nccl.__file__), "include")]
    except Exception:
        return []


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), out: str | None = None) -> str:
    """extra_flags/out: tuning variants (e.g. -DMB_H_MINB=3) built next to the
    default library for A/B runs on the GPU box (MOLOCH_B200_LIB selects one)."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build" if out is None else "build_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    inc = _nccl_include()

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = ([nvcc] + NVCC_FLAGS + list(extra_flags) + inc + (["-Xptxas", "-v"] if verbose else []) +
               ["-c", os.path.join(CSRC, src), "-o", obj])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    target = LIB if out is None else out
    # -Bsymbolic: the library's own calls bind to its own definitions even when another build of the same
    # sources (the fast-mode library, a tuning variant) is loaded into the same process
    cmd = [nvcc, "-shared", "-Xlinker", "-Bsymbolic", "-o", target] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return target


def build_fast(force: bool = False) -> str:
    """libmoloch_b200_fast.so: -fmad=true instead of -fmad=false, everything else identical."""
    if not force and os.path.exists(LIB_FAST):
        t = os.path.getmtime(LIB_FAST)
        deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
        if not any(os.path.getmtime(d) > t for d in deps):
            return LIB_FAST
    keep = list(NVCC_FLAGS)
    try:
        NVCC_FLAGS[NVCC_FLAGS.index("-fmad=false")] = "-fmad=true"
        return build_library(out=LIB_FAST)
    finally:
        NVCC_FLAGS[:] = keep


if __name__ == "__main__":
    import sys
    print(build_library(force="-f" in sys.argv, verbose="-v" in sys.argv))
    if "--fast" in sys.argv:
        print(build_fast(force="-f" in sys.argv))
