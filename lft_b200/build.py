"""Build the C-ABI CUDA library in-tree: lft_b200/_lib/liblft_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# LFT_VARIANT=<name> builds / loads an experimental variant next to the product library (lft_b200/_lib_<name>/), compiled
# with the extra flags in LFT_DEFINES (e.g. "-DLFT_EXPERIMENT_X"); a variant that exists is loaded as is.  A/B timing only.
VARIANT = os.environ.get("LFT_VARIANT", "")
OUT_DIR = os.path.join(HERE, "_lib" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(OUT_DIR, "liblft_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
FLAGS += os.environ.get("LFT_DEFINES", "").split()
if os.environ.get("LFT_TIMELINE"):
    FLAGS.append("-DLFT_TIMELINE")
if os.environ.get("LFT_EXPERIMENT_NOSTORE"):    # timing experiment (wrong results): k_spa_embed_qkv* never store
    FLAGS.append("-DLFT_EXPERIMENT_NOSTORE")
if os.environ.get("LFT_EXPERIMENT_NOSTREAM"):   # timing experiment (wrong results): weight slabs are never copied
    FLAGS.append("-DLFT_EXPERIMENT_NOSTREAM")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


LAST_BUILD_COMPILED = False   # True once build() has actually run nvcc in this process (build_mode evidence)


def build(force: bool = False, verbose: bool = False) -> str:
    global LAST_BUILD_COMPILED
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and (VARIANT or open(stamp).read() == dig):
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build liblft_b200.so")
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".ptxas.log")
        open(log, "w").write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(dig)
    LAST_BUILD_COMPILED = True
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
