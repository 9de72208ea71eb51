"""Build libivs.so (the C-ABI CUDA handler) in-tree for sm_100a.

    python -m inovesa_b200.build          # -> inovesa_b200/lib/libivs.so

nvcc cross-compiles without a GPU.  Device code is compiled with -fmad=false (no implicit
contraction, see csrc/ivs_common.cuh), host set-up code with -ffp-contract=off.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libivs.so")

CU = ["ivs_api.cu", "ivs_maps.cu", "ivs_reduce.cu", "ivs_field.cu", "ivs_track.cu", "ivs_dynrf.cu", "ivs_train.cu",
      "ivs_step.cu", "ivs_fused.cu", "ivs_stream8.cu", "ivs_resident.cu"]
CPP = ["ivs_host.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(HERE, "..", "include", f) for f in ("ivs.h", "ivs_host.h")]
    objs = []
    for f in CU:
        src, obj = os.path.join(CSRC, f), os.path.join(objdir, f + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            cmd = [NVCC, "-ccbin", "/usr/bin/g++", "-std=c++17", "-O3", "-lineinfo", "-fmad=false", *ARCH,
                   "-Xcompiler", "-fPIC,-ffp-contract=off", "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    for f in CPP:
        src, obj = os.path.join(CSRC, f), os.path.join(objdir, f + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC",
                   "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _newer(LIB, objs):
        cmd = [NVCC, "-ccbin", "/usr/bin/g++", "-shared", *ARCH, "-o", LIB, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
