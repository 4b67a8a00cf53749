"""Builds platypus_b200/libplatypus_b200.so (sm_100a only) with nvcc.  In-tree so the .so
travels to the GPU box with the snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "libplatypus_b200.so")
SRCS = ["plb_api.cu"]
DEPS = ["plb_api.cu", "plb_kernels.cuh", "plb_dp.cuh", os.path.join("..", "..", "include", "platypus_b200.h"), "plb_kmer.cuh", "plb_select.cuh", "plb_stage.cuh", "plb_synth.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
              "-shared", "-Xcompiler", "-fPIC,-fopenmp", "-lgomp"]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(HERE, d)) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("PLB_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRCS
    r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
