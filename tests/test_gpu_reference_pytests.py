"""The reference's OWN Python tests, unmodified, against the GPU backend.

`__graft_entry__.build()` stages /root/reference/test/test_{multilinear,multicubic,nearest}_{regular,rectilinear}.py and
test_interpn.py byte for byte under baseline/_ref/test/ (git-ignored: reference files are never committed; the directory
travels to the GPU box with the snapshot like any other built artefact). They `import interpn`; tests/shim/interpn is an
alias package that re-exports interpn_b200, so every `interpn.raw.*` call, every pydantic class method
(`new / eval / check_bounds / model_dump_json`) and `interpn.interpn(...)` in those files runs on the GPU through the C ABI.
Run once per arithmetic flavour. Needs a B200: `pytest -m gpu`.
"""

import hashlib
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "test")
SHIM = os.path.join(ROOT, "tests", "shim")
FILES = ["test_interpn.py"] + [f"test_{m}_{k}.py" for m in ("multilinear", "multicubic", "nearest") for k in ("regular", "rectilinear")]


def _staged():
    return all(os.path.exists(os.path.join(STAGED, f)) for f in FILES)


@pytest.mark.parametrize("flavour", ["fma", "strict"])
def test_reference_python_tests_pass_unmodified_on_the_gpu_backend(flavour):
    if not _staged():
        pytest.skip("reference tests not staged (run __graft_entry__.build() where /root/reference exists)")
    if os.path.isdir("/root/reference/test"):  # in the build container: prove the staged copies are unmodified
        for f in FILES:
            a = hashlib.sha256(open(os.path.join(STAGED, f), "rb").read()).hexdigest()
            b = hashlib.sha256(open(os.path.join("/root/reference/test", f), "rb").read()).hexdigest()
            assert a == b, f"{f} differs from the reference's file"
    env = dict(os.environ, INTERPN_B200_ARITHMETIC=flavour, PYTHONPATH=os.pathsep.join([SHIM, ROOT]))
    env.pop("INTERPN_B200_LIBRARY", None)
    # a conftest-free run rooted at the staged directory: nothing of this repo's test configuration leaks in
    probe = (
        "import interpn, interpn_b200, sys; "
        "assert interpn.raw.interpn_linear_regular_f64 is interpn_b200.raw.interpn_linear_regular_f64; "
        f"assert interpn_b200._lib.ARITHMETIC == '{flavour}'; print(interpn.__backend__)"
    )
    subprocess.run([sys.executable, "-c", probe], cwd=STAGED, env=env, check=True, timeout=300)
    count = (
        "import interpn_b200, atexit; n0 = interpn_b200.launch_count(); "
        "atexit.register(lambda: print('GPU_LAUNCHES', interpn_b200.launch_count() - n0))"
    )
    plugin = os.path.join(STAGED, "..", "_launch_probe.py")
    with open(plugin, "w") as fh:
        fh.write(count.replace("; ", "\n") + "\n")
    env["PYTHONPATH"] = os.pathsep.join([SHIM, ROOT, os.path.dirname(plugin)])
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-p", "_launch_probe", "--rootdir", STAGED, "-c", os.devnull, *FILES],
        cwd=STAGED, env=env, capture_output=True, text=True, timeout=1200,
    )  # fmt: skip
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-30:])
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
    launches = [int(line.split()[1]) for line in r.stdout.splitlines() if line.startswith("GPU_LAUNCHES")]
    assert launches and launches[0] > 50, tail  # the reference's tests really ran on the GPU
