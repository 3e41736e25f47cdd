"""Build libnvsm_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m cunvsm_b200.build            # build if stale
    python -m cunvsm_b200.build --force
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnvsm_b200.so")
SOURCES = ["nvsm.cu"]
DEPS = ["nvsm.cu", "kernels.cuh", "common.cuh", "gemm_simt.cuh", "gemm_tcgen05.cuh", "gemm_tcgen05_2cta.cuh", "similarity.cuh", "peer_allreduce.cuh", "pull_update.cuh", "score_ring.cuh", "sampler.cuh", "nccl_dyn.h", "microbench.cuh", "ops.cuh", "ops_host.inc",
        os.path.join("..", "..", "include", "nvsm_b200.h")]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.sep not in p or os.path.exists(p)):
            return p
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS if os.path.exists(os.path.join(CSRC, d)))


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-ccbin", host_cxx, "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
           "-Xptxas", "-v" if verbose else "-O3",
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libnvsm_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
