"""The wide BVHs on the CPU: the greedy 4-wide collapse, the 8-wide quantised collapse (csrc/bvh8.cuh) and the traversal code
(csrc/bvh.cuh: trav_run, trav_run4, trav_run8) are compiled for the host (tests/host/cuda_host_shim.h stands in for the device
intrinsics) and checked against a brute-force loop over all triangles: hit, t and barycentrics bit-identical for random,
axis-parallel, grazing, far-away and surface-start rays; the 8-wide traversal run in quanta of 8 steps agrees as well."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("bvh8") / "bvh8_host"
    src = ROOT / "tests" / "host" / "bvh8_host.cpp"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I/usr/local/cuda/include", f"-I{src.parent}", str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.mark.parametrize("tris,rays,seed", [(4000, 12000, 1), (17, 3000, 5), (300, 6000, 3), (12000, 6000, 11)])
def test_bvh8_collapse_and_traversal_match_brute_force(harness, tris, rays, seed):
    r = subprocess.run([str(harness), str(tris), str(rays), str(seed)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-1000:]
