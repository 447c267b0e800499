"""Build libb2c.so (the C-ABI shared library with the sm_100a kernels) in-tree with nvcc.

Strict IEEE binary32 everywhere (the path must reproduce Java float semantics):
-fmad=false (no FMA contraction), -prec-div/-prec-sqrt=true, -ftz=false on the device;
-ffp-contract=off for the host code (BVH build, shape registration).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb2c.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
    "-shared", "-cudart", "shared",
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(os.path.dirname(HERE), "include", "b2c.h"))
    d.append(os.path.abspath(__file__))
    return d


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in deps())


def build(force=False, verbose=False, out=None, extra=()):
    """extra / B2C_NVCC_EXTRA: additional nvcc flags for tuning experiments (e.g. -DGJK_MINB=5), never used by default."""
    out = out or OUT
    extra = list(extra) + os.environ.get("B2C_NVCC_EXTRA", "").split()
    if not force and not extra and out == OUT and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libb2c.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
