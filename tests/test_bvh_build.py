"""Host-side BVH builder (csrc/bvh_build.h): structural invariants, parallel == sequential."""
import os
import subprocess

import pytest

from helpers import ROOT


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("bvh") / "bvh_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "raytracingoneweekendapplication_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "bvh_check.cpp")])
    return exe


@pytest.mark.parametrize("n,threads", [(0, 0), (1, 0), (2, 0), (3, 1), (37, 0), (5000, 1), (20000, 0), (20000, 3)])
def test_bvh_invariants(checker, n, threads):
    out = subprocess.run([checker, str(n), str(threads)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().splitlines()[-1].startswith("OK"), out.stdout + out.stderr
    assert "wide8" in out.stdout and "wide4" in out.stdout   # the collapsed wide trees were checked too (csrc/bvh_wide.h)
