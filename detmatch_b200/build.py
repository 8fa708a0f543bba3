"""Builds detmatch_b200/lib/libpcfe.so (the C-ABI library of include/pcfe.h) with nvcc for sm_100a.

In-tree build: the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libpcfe.so")
SOURCES = ["voxelize.cu", "hv_global.cu", "hv_bucket.cu", "points_in_boxes.cu", "roiaware_pool3d.cu", "scatter.cu"]
HEADERS = [os.path.join(CSRC, "pcfe_common.cuh"), os.path.join(CSRC, "hv_common.cuh"), os.path.join(CSRC, "hv_cluster.cuh"), os.path.join(CSRC, "pib_dev.cuh"),
           os.path.join(ROOT, "include", "pcfe.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",            # no implicit FMA contraction anywhere: bit-exactness is the contract
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """`defines`/`out` build a tuning variant (e.g. defines=["PCFE_BUCKET_THREADS=128"],
    out="lib/libpcfe_bt128.so"); the default build takes neither."""
    target = os.path.join(HERE, out) if out else LIB
    if not force and not out and not stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    cmd += ["-D" + d for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", target]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpcfe.so")
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[3:] for a in sys.argv[1:] if a.startswith("-o=")]
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
