"""TEST INFRASTRUCTURE ONLY.  Compiles the libsdtree sources with plain g++ and
-DSDT_HOSTEMU (kernels become serial loops, CUDA runtime calls become malloc/memcpy; see
csrc/sdt_platform.h) into tests/hostemu/libsdtree_hostemu.so, so that the `-m "not gpu"`
tests can hold the kernels' index logic against the oracle in a container without a GPU.
The package never loads this file and nothing is benchmarked through it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "practical_path_guiding_lab_b200", "csrc")
OUT = os.path.join(HERE, "libsdtree_hostemu.so")


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "sdtree.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DSDT_HOSTEMU", "-x", "c++",
           os.path.join(CSRC, "sdtree.cu"), "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("hostemu build failed:\n" + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
