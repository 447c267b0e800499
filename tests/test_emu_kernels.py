"""CPU: the device code of narrowphase.cuh / compound.cuh / raycast.cuh / convexcast.cuh (with gjk.cuh, epa.cuh) compiled
for the host with a shim of the CUDA built-ins (tests/emu/) and run against the oracle: every child manifold, raw record and
ray hit must be bit-identical.  Test infrastructure only — nothing here is a product path; the -m gpu tests remain the parity
tests proper."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
FLAGS = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-DB2C_HOST_EMULATION", "-I", EMU, "-w"]


@pytest.fixture(scope="module")
def binaries(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu")
    bins = {}
    for name in ("emu_compound", "emu_ray", "emu_narrowphase"):
        exe = str(out / name)
        subprocess.check_call(FLAGS + ["-o", exe, os.path.join(EMU, name + ".cpp")], cwd=EMU)
        bins[name] = exe
    return bins


@pytest.mark.parametrize("args", [("150", "1.0"), ("200", "0.45"), ("120", "0.8", "mesh"), ("150", "0.7", "mesh", "-0.5")])
def test_compound_kernels_match_the_oracle_on_the_host(binaries, args):
    """k_compound_expand / k_compound_gjk / k_epa<2>,<1> / k_compound_manifold / k_compound_mesh over 8 steps with bodies going
    to sleep and waking up: child manifold headers, points, raw records and counters, bit for bit."""
    r = subprocess.run([binaries["emu_compound"], *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("args", [("200", "3000"), ("400", "3000", "tilt"), ("300", "2000", "tilt", "mask")])
def test_ray_kernels_match_the_oracle_on_the_host(binaries, args):
    """k_ray_aabbs / k_ray_chunks / k_ray_test against convex bodies, a rotated triangle mesh, a tilted static plane and
    compounds, with and without a callback filter: hit body, fraction, normal and point, bit for bit.  Then k_convex_sweep
    (convexcast.cuh) on the same scene: sphere / box / hull casts with random bases against convex bodies, the mesh and
    compounds, plus the sweeps that must land in the reference's throwing static-plane branch."""
    r = subprocess.run([binaries["emu_ray"], *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("args", [("200", "0.8"), ("300", "0.5")])
def test_narrowphase_kernels_match_the_oracle_on_the_host(binaries, args):
    """k_carry / k_sphere_sphere / k_convex_plane / k_gjk / k_mesh_query / k_gjk_tri / k_epa<2>,<1> / k_manifold_cc /
    k_mesh_manifold over 8 steps of boxes, spheres and hulls on a plane and a triangle mesh: raw detector records (method and
    iteration count included), manifold headers and points, bit for bit."""
    r = subprocess.run([binaries["emu_narrowphase"], *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
