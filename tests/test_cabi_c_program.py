"""The boundary used from plain C: tests/cabi_demo.c includes include/interpn_b200.h, links libinterpn_b200.so and checks
known answers (the reference's own linear-field tests), error messages and the failure-index rule — no Python between
the caller and the C ABI.

* CPU (`-m "not gpu"`): the program compiles against the header with -Wall -Werror, links against the library (every
  symbol it uses resolves), reports argument errors with the reference's messages and, without an sm_100 device, stops
  with the library's "no device" status (exit code 77) instead of computing anything on the CPU.
* GPU (`-m gpu`): it runs to completion, in both arithmetic flavours.
"""

import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "interpn_b200")


def build(tmp_path, libname):
    exe = str(tmp_path / ("cabi_demo_" + libname))
    lib = os.path.join(PKG, f"lib{libname}.so")
    assert os.path.exists(lib), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "cabi_demo.c"),
           "-o", exe, "-L", PKG, f"-l{libname}", "-lm", f"-Wl,-rpath,{PKG}"]  # fmt: skip
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_c_program_compiles_links_and_refuses_without_a_device(tmp_path):
    import torch

    exe = build(tmp_path, "interpn_b200")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)  # argument checks passed, then "no device"
        assert "no sm_100 device" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("libname", ["interpn_b200", "interpn_b200_fma"])
def test_c_program_known_answers_on_the_gpu(tmp_path, libname):
    exe = build(tmp_path, libname)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
    assert f"arithmetic flavour {1 if libname.endswith('fma') else 0}" in r.stdout
