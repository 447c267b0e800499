"""The C++ host mirror (include/b2c_host.hpp): compiles and links against libb2c.so on the CPU box; on the GPU
it runs the reference's call sequence and produces the hand-derived stack contacts."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "libgdx-jbullet_b200")
EXE = os.path.join(ROOT, "tests", "_host_mirror_demo")


def build_demo():
    import __graft_entry__ as ge
    ge.build()
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host_mirror_demo.cpp"),
           "-L", PKG, "-lb2c", "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", EXE]
    subprocess.check_call(cmd)


def test_host_mirror_compiles_and_links():
    build_demo()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_mirror_runs_reference_call_sequence(gpu_pkg):
    build_demo()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "pairs 3 manifolds 3"
    assert lines[1:4] == ["pair 1 2", "pair 2 3", "pair 3 4"]
    # ground (uid 1) vs lowest box (uid 2): normal on B points from the box (B) down to the ground (A)
    assert "contacts 1 normal 0.000 -1.000 0.000" in lines[4]
    # first step: all three pairs are new; the three stacked boxes form one island (tag = smallest body index), ground is static
    assert lines[5] == "deltas +3 -0 islands 1 tags -1 1 1 1"
    # ray and convex sweep straight down onto the stack (top of the third box at y = 6): both hit body 4; the sphere of radius
    # 0.25 touches after (10 - 6.25) / 20 of the way
    assert lines[6].startswith("ray uid 4 y 6.00 | sweep uid 4 fraction 0.18") and lines[6].endswith("normal 0.00 1.00 0.00"), lines[6]
    # the drop-in sequence and the fast path, stepped side by side over a drifting scene, tell the same
    # story (pair list == mirror maintained from the deltas, same touching manifolds, bit-identical contact points)
    steps = [l for l in lines if l.startswith("step ")]
    assert len(steps) == 5 and not any(l.startswith("FAIL") for l in lines), out.stdout
    counts = [int(l.split()[3]) for l in steps]
    assert min(counts) > 150
    assert any(int(l.split()[5]) < 0 for l in steps[1:]) and any(int(l.split()[4]) > 0 for l in steps[1:]), "pairs must come and go"
    assert lines[-1].startswith("broadphase aabb -1e+30 1e+30")
