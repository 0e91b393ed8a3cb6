"""Parity pinned on OUTPUTS OF THE REFERENCE ITSELF (tests/golden/ref_docs.npz).

The fixture holds 34 660 points evaluated by the published reference wheel (an `fma`-feature build of the crate,
/root/reference/pyproject.toml:72) together with their inputs, extracted from the Plotly JSON of the reference's
documentation figures by tests/golden/extract_ref_docs.py: 1-D and 2-D multicubic (regular + rectilinear,
linearize_extrapolation=False, interior + extrapolation) and 2-D rectilinear nearest.

* CPU: the oracle in fma mode must reproduce every point bit for bit; in strict mode (the crate's default features,
  same code with the mul_add sites unfused) it must stay within the tolerance BASELINE.json's north_star states for
  f64: 4 ULP or 1e-13 relative.
* GPU (`-m gpu`): libinterpn_b200_fma.so, called through the C ABI via interpn_b200.raw in a child process (the flavour
  is a per-process choice), must reproduce every point bit for bit; the strict library within the same tolerance.
"""

import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PATH = os.path.join(HERE, "golden", "ref_docs.npz")
G = np.load(PATH)
META = json.loads(bytes(G["meta"]).decode())
NAMES = [m["name"] for m in META]
BY_NAME = {m["name"]: m for m in META}

REL_TOL = 1e-13  # north_star: "f64 results within 4 ULP or 1e-13 relative"
ULP_TOL = 4


def case(name):
    m = BY_NAME[name]
    nd = m["ndims"]
    c = dict(m, vals=G[f"{name}/vals"], obs=[G[f"{name}/obs{d}"] for d in range(nd)], out=G[f"{name}/out"])
    if m["rect"]:
        c["grids"] = [G[f"{name}/grid{d}"] for d in range(nd)]
    else:
        c["dims"] = [int(v) for v in G[f"{name}/dims"]]
        c["starts"], c["steps"] = G[f"{name}/starts"], G[f"{name}/steps"]
    return c


def bits_equal(a, b):
    return a.shape == b.shape and bool(np.all(a.view(np.uint64) == b.view(np.uint64)))


def within_tolerance(got, want):
    """4 ULP or 1e-13 relative, per point (the scale of one 2-D evaluation is the largest value in play, so the
    relative test is taken against max(|want|, |data|) like the reference's own approx checks)."""
    ulp = np.spacing(np.maximum(np.abs(want), np.finfo(np.float64).tiny))
    err = np.abs(got - want)
    return bool(np.all((err <= ULP_TOL * ulp) | (err <= REL_TOL * np.maximum(np.abs(want), 1.0))))


def oracle_eval(oracle, c, fma):
    if c["rect"]:
        return oracle.interpn_rectilinear(c["method"], c["grids"], c["vals"], c["obs"], linearize_extrapolation=c["linearize"], fma=fma)
    return oracle.interpn_regular(c["method"], c["dims"], c["starts"], c["steps"], c["vals"], c["obs"],
                                  linearize_extrapolation=c["linearize"], fma=fma)  # fmt: skip


def test_fixture_is_the_reference_figures():
    assert len(META) == 9
    assert sum(G[f"{n}/out"].size for n in NAMES) == 34660
    assert {m["method"] for m in META} == {"cubic", "nearest"}
    # extrapolation is part of the pinned slice: queries leave the grid on both sides
    c = case("cubic1d_regular_sine")
    assert c["obs"][0].min() < c["starts"][0] and c["obs"][0].max() > c["starts"][0] + c["steps"][0] * (c["dims"][0] - 1)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_fma_is_bit_identical_to_the_reference_wheel(oracle, name):
    c = case(name)
    got = oracle_eval(oracle, c, fma=True)
    assert bits_equal(got, c["out"]), f"{name}: {(got.view(np.uint64) != c['out'].view(np.uint64)).sum()} of {got.size} points differ"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_strict_is_within_the_stated_tolerance(oracle, name):
    c = case(name)
    got = oracle_eval(oracle, c, fma=False)
    assert within_tolerance(got, c["out"]), name
    if c["method"] == "nearest":
        assert bits_equal(got, c["out"])  # index path: bit-exact in either flavour


_CHILD = r"""
import json, sys
import numpy as np
import interpn_b200 as ib
from interpn_b200 import _lib
G = np.load(sys.argv[1]); meta = json.loads(bytes(G["meta"]).decode()); res = {}
assert _lib.lib.interpn_b200_arithmetic() == (1 if sys.argv[3] == "fma" else 0)
n0 = ib.launch_count()
for m in meta:
    n, nd = m["name"], m["ndims"]
    obs = [G[f"{n}/obs{d}"] for d in range(nd)]; out = np.zeros_like(G[f"{n}/out"])
    extra = (m["linearize"],) if m["method"] == "cubic" else ()
    if m["rect"]:
        getattr(ib.raw, f"interpn_{m['method']}_rectilinear_f64")([G[f"{n}/grid{d}"] for d in range(nd)], G[f"{n}/vals"], *extra, obs, out)
    else:
        getattr(ib.raw, f"interpn_{m['method']}_regular_f64")([int(v) for v in G[f"{n}/dims"]], G[f"{n}/starts"], G[f"{n}/steps"], G[f"{n}/vals"], *extra, obs, out)
    res[n] = out
assert ib.launch_count() - n0 >= len(meta)
np.savez(sys.argv[2], **res)
"""


def _gpu_outputs(flavour):
    with tempfile.TemporaryDirectory() as d:
        dst = os.path.join(d, "out.npz")
        env = dict(os.environ, INTERPN_B200_ARITHMETIC=flavour)
        env.pop("INTERPN_B200_LIBRARY", None)
        r = subprocess.run([sys.executable, "-c", _CHILD, PATH, dst, flavour], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
        z = np.load(dst)
        return {n: z[n] for n in NAMES}


@pytest.mark.gpu
def test_cuda_fma_library_is_bit_identical_to_the_reference_wheel():
    got = _gpu_outputs("fma")
    bad = {n: int((got[n].view(np.uint64) != G[f"{n}/out"].view(np.uint64)).sum()) for n in NAMES}
    assert all(v == 0 for v in bad.values()), bad


@pytest.mark.gpu
def test_cuda_strict_library_is_within_the_stated_tolerance_of_the_reference_wheel():
    got = _gpu_outputs("strict")
    for n in NAMES:
        assert within_tolerance(got[n], G[f"{n}/out"]), n
    assert bits_equal(got["nearest2d_rectilinear"], G["nearest2d_rectilinear/out"])
