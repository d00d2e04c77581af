"""Builds libsdtree.so in-tree with nvcc for sm_100a (B200).

    python -m practical_path_guiding_lab_b200.build [--force]

Flags: IEEE arithmetic only -- no -use_fast_math, and -fmad=false so that every fp32
operation is separately rounded (leaf / node indices and the refine topology must be
bit-exact against the oracle, which numpy evaluates without fused multiply-add).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsdtree.so")
SOURCES = ["sdtree.cu"]
DEPS = ["sdtree.cu", "sdt_platform.h", "sdt_core.h", "sdt_exec.h", "sdt_impl.h", "sdt_query.inl",
        "sdt_splat.inl", "sdt_refine.inl", "sdt_io.inl", "sdt_nccl.inl", os.path.join("..", "..", "include", "sdtree.h")]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libsdtree.so cannot be built (there is no CPU fallback)")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-o", OUT]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
