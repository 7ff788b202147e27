"""The device arithmetic header compiled for the host (g++) against the oracle's C restatement:
catches kernel-side solver bugs on a machine without a GPU.  Test tooling only."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_device_solver_matches_oracle(tmp_path):
    exe = str(tmp_path / "host_harness")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.check_call(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-std=c++17",
                           "-I", ROOT, os.path.join(ROOT, "tests", "host_harness.cpp"), "-o", exe, "-lm"], env=env)
    out = subprocess.run([exe, "40000"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
