"""The `fma` flavour of the library (libinterpn_b200_fma.so: the arithmetic of the reference crate built with
`--features fma`, which is how the reference's Python wheel is built, pyproject.toml:72) against the oracle in fma mode.

The flavour is chosen once per process (INTERPN_B200_ARITHMETIC, read by interpn_b200/_lib.py and by
oracle/oracle.py), so the parity suite and the reference's own test suite are re-run in a child process with the
variable set: every assertion of tests/test_gpu_parity.py and tests/test_gpu_reference_suite.py then compares the fma
library with the fma oracle, bit for bit. Needs a B200: `pytest -m gpu`."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    env = dict(os.environ, INTERPN_B200_ARITHMETIC="fma")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", *args],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)  # fmt: skip
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in r.stdout, tail


def test_child_process_really_loads_the_fma_library():
    code = (
        "import interpn_b200._lib as L, ctypes; "
        "assert L.LIB_PATH.endswith('libinterpn_b200_fma.so'), L.LIB_PATH; "
        "assert L.lib.interpn_b200_arithmetic() == 1; "
        "from oracle import oracle; assert oracle.DEFAULT_FMA"
    )
    env = dict(os.environ, INTERPN_B200_ARITHMETIC="fma")
    subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, check=True)


def test_fma_and_strict_flavours_differ_where_the_reference_builds_differ():
    """A smoke check that the flavour is not a no-op: on smooth data the two libraries agree to a few ulp
    (CHANGELOG.md:114-118) but not bit for bit."""
    code = (
        "import numpy as np, interpn_b200 as ib\n"
        "rng = np.random.default_rng(3); dims=[12]*3; n=20000\n"
        "starts=np.zeros(3); steps=np.full(3, 0.37); vals=rng.standard_normal(12**3)\n"
        "obs=[rng.uniform(0.5, 3.5, n) for _ in range(3)]; out=np.zeros(n)\n"
        "ib.raw.interpn_cubic_regular_f64(dims, starts, steps, vals, True, obs, out)\n"
        "np.save(__import__('sys').argv[1], out)\n"
    )
    import tempfile

    import numpy as np

    outs = {}
    with tempfile.TemporaryDirectory() as d:
        for flavour in ("strict", "fma"):
            path = os.path.join(d, flavour + ".npy")
            env = dict(os.environ, INTERPN_B200_ARITHMETIC=flavour)
            subprocess.run([sys.executable, "-c", code, path], cwd=ROOT, env=env, check=True)
            outs[flavour] = np.load(path)
    a, b = outs["strict"], outs["fma"]
    assert not np.array_equal(a, b)
    assert np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-3)) < 1e-12


def test_parity_suite_in_fma_flavour():
    _run(["tests/test_gpu_parity.py"])


def test_reference_suite_in_fma_flavour():
    _run(["tests/test_gpu_reference_suite.py"])
