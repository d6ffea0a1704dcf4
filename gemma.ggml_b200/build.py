"""Build the two native libraries IN-TREE (they travel to the GPU box with the snapshot):

  libggb_cuda.so  — the thin C-ABI CUDA layer (csrc/*.cu), sm_100a only
  libggb_host.so  — the host C++ side (host/*.cpp): hpc mul_mat dispatch, graph builder,
                    tensor_dump; links against libggb_cuda.so

Usage: python gemma.ggml_b200/build.py [--force]
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OBJ = os.path.join(HERE, "build")
CUDA_LIB = os.path.join(HERE, "libggb_cuda.so")
HOST_LIB = os.path.join(HERE, "libggb_host.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", *os.environ.get("GGB_NVCC_EXTRA", "").split(),
    "-fmad=false",                 # no implicit FMA contraction: arithmetic matches the oracle
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-I", os.path.join(ROOT, "include"),
]
NCCL_INC = "/usr/include"
for cand in ("/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/include",):
    if os.path.exists(os.path.join(cand, "nccl.h")):
        NCCL_INC = cand
        break
CXX = os.environ.get("CXX", "g++")
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall",
             "-I", os.path.join(ROOT, "include"), "-I", HOST]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + os.path.basename(cmd[-1]))
    if r.stderr.strip():
        sys.stderr.write(r.stderr)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + \
        glob.glob(os.path.join(HOST, "*.h")) + [os.path.abspath(__file__)]
    # ---- CUDA layer
    objs = []
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            extra = ["-Xptxas", "-v"] if verbose else []
            _run([NVCC] + NVCC_FLAGS + extra + ["-I", NCCL_INC, "-c", src, "-o", obj])
    if force or _stale(CUDA_LIB, objs):
        _run([NVCC, "-shared", "-o", CUDA_LIB] + objs + ["-lcudart", "-ldl"])
    # ---- host C++ layer
    hobjs = []
    for src in sorted(glob.glob(os.path.join(HOST, "*.cpp"))):
        obj = os.path.join(OBJ, "host_" + os.path.basename(src)[:-4] + ".o")
        hobjs.append(obj)
        if force or _stale(obj, [src] + headers):
            _run([CXX] + CXX_FLAGS + ["-c", src, "-o", obj])
    if hobjs and (force or _stale(HOST_LIB, hobjs + [CUDA_LIB])):
        _run([CXX, "-shared", "-o", HOST_LIB] + hobjs +
             ["-L", HERE, "-lggb_cuda", "-Wl,-rpath,$ORIGIN", "-lpthread"])
    return CUDA_LIB, HOST_LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", CUDA_LIB, HOST_LIB)
