"""Build libpomcp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pomcp_jl_b200.build [--force] [--verbose]

-fmad=false: every fp64 expression keeps the IEEE operation order written in the source, so the
device results can be compared bit for bit with the CPU oracle (see csrc/detmath.cuh).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libpomcp_b200.so")
SOURCES = ["abi.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh"))  # every header abi.cu can include


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "pomcp_b200.h"),
                                                                 os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-fmad=false", "-ccbin", host_cxx, "-shared", "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
           "-o", SO] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(r.stderr)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(SO)
