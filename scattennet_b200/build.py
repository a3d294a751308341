"""Build ``libscatt.so`` in-tree with nvcc for sm_100a (no torch dependency).

``python -m scattennet_b200.build`` or ``scattennet_b200.build.build()``.
nvcc cross-compiles without a GPU; the ``.so`` is git-ignored but travels to
the GPU box with the repo snapshot.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libscatt.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")
SOURCES = ["api.cu", "rowwise.cu", "frontend_tc.cu", "prefetch.cu", "gemm_simt.cu", "gemm_tc.cu", "block_tc.cu", "attention.cu", "attention_tc.cu", "attention_fa.cu", "fusion_tc.cu", "lstm.cu", "heads.cu", "ctc.cu", "peer.cu"]
HEADERS = ["common.cuh", "tc_ptx.cuh", "tc_epi.cuh", "tc_host.cuh", os.path.join("..", "..", "include", "scatt.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libscatt.so cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = OUT + ".digest"
    return os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read().strip() == _digest()


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = OUT) -> str:
    """Compile every CUDA source for sm_100a and link ``libscatt.so``; returns its path.
    ``extra_flags`` / ``out`` build instrumented variants for the dev tools."""
    if not force and not extra_flags and is_current():
        return OUT
    nvcc = _nvcc()
    obj_dir = OBJ_DIR if not extra_flags else OBJ_DIR + "_" + hashlib.sha1(" ".join(extra_flags).encode()).hexdigest()[:8]
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        with open(obj + ".ptxas.log", "w") as fh:
            fh.write(res.stderr)
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    link = [nvcc, *ARCH, "-shared", "-o", out, *objs, "-cudart", "static", "-Xcompiler", "-fPIC"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    if out == OUT:
        with open(OUT + ".digest", "w") as fh:
            fh.write(_digest())
    return out


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
