"""Builds the C-ABI shared library for sm_100a with nvcc (in-tree, so the .so
travels to the GPU box with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
# S256_LIB=<path> selects a prebuilt variant (tuning experiments); it is never rebuilt.
LIB = os.environ.get("S256_LIB") or os.path.join(LIBDIR, "libsecp256k1_b200.so")
SOURCES = ["api.cu", "kern_ct.cu", "codecs.cpp"]
HEADERS = ["fe.cuh", "sc.cuh", "point.cuh", "sha256.cuh", "kernels.cuh", "microbench.cuh", "launchers.h", "msm.cuh", "vm.cuh", "fe_sqr_gen.cuh"]


def nvcc_cmd(extra=(), out=None):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
            "-Xptxas", "-v", *extra,
            "-o", out or LIB] + [os.path.join(CSRC, s) for s in SOURCES]


def is_stale():
    if os.environ.get("S256_LIB"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "secp256k1_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=(), out=None):
    if out is None and not force and not is_stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    if out:
        os.makedirs(os.path.dirname(out), exist_ok=True)
    p = subprocess.run(nvcc_cmd(extra, out), capture_output=True, text=True)
    log = p.stdout + p.stderr
    with open((out or os.path.join(LIBDIR, "build")) + ".log" if out else os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write(log)
    if p.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libsecp256k1_b200.so")
    if verbose:
        print(log)
    return out or LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
