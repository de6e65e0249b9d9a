"""Builds libwifi_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libwifi_b200.so")
SRC = os.path.join(HERE, "csrc", "wifi_b200.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("wifi_b200.cu", "wifi_common.cuh", "viterbi.cuh", "rx_kernels.cuh", "tx_kernels.cuh")] + [
    os.path.join(HERE, "..", "include", f) for f in ("wifi_b200.h", "wifi_detmath.h")]
# -fmad=false: no implicit FMA contraction -- the numerical contract of include/wifi_detmath.h
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


HASH = SO + ".srchash"


def _src_hash():
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in DEPS:
        h.update(open(d, "rb").read())
    return h.hexdigest()


def stale():
    """Stale = the sources' content differs from what the binary was built from (mtimes are not
    trusted: the tree is copied to the GPU box)."""
    if not os.path.exists(SO):
        return True
    if not all(os.path.exists(d) for d in DEPS):
        return False  # sources stripped: use the shipped binary
    try:
        return open(HASH).read().strip() != _src_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC]
    subprocess.check_call(cmd, cwd=HERE)
    with open(HASH, "w") as f:
        f.write(_src_hash())
    return SO


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
