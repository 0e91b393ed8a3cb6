"""CPU re-check of the division-avoiding sequences the CUDA kernels use (device_math.cuh:
markstein_div / exact_div, fast_cell, nearest_upper) against the IEEE divisions of the reference
(multilinear/regular.rs:414-425, nearest/regular.rs:259-293). The sequences are restated with host
operations in tests/numeric_tricks.cpp; the GPU parity suite then checks the device code itself."""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_sequences_match_ieee_division(tmp_path):
    exe = tmp_path / "numeric_tricks"
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(HERE, "numeric_tricks.cpp")],
                   check=True, env=env)  # fmt: skip
    r = subprocess.run([str(exe), "30"], capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
    trials, proven = (int(v) for v in r.stdout.split()[1:3])
    assert proven > 0.8 * trials  # the fast path must actually carry the load it is checked on
