"""In-tree build of libnesosim_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnesosim_b200.so")
# The CUDA runtime is linked as a shared library: inside a PyTorch process the library then binds to the runtime torch
# has already loaded (one runtime per process); standalone, the rpath finds the toolkit's own copy.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "--cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
              "-shared", "-Xcompiler", "-fPIC,-pthread", "-Xptxas", "-v"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    hdr = os.path.join(os.path.dirname(HERE), "include", "nesosim_b200.h")
    return any(os.path.getmtime(p) > t for p in sources() + [hdr])


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB, os.path.join(CSRC, "nesosim_abi.cu")]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed (%d); see %s" % (res.returncode, log))
    return LIB


if __name__ == "__main__":
    build(force=True)
