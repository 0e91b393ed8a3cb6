"""CPU exercise of the copy-thread pool that stages pageable caller memory for the host executor
(interpn_b200/csrc/copy_pool.h, used by host_exec.cuh): tests/copy_pool_test.cpp compiled with g++ and run, with
non-temporal stores on and off and with a pool of zero, three and the default number of threads."""

import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    path = tmp_path_factory.mktemp("copy_pool") / "copy_pool_test"
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", str(path), os.path.join(HERE, "copy_pool_test.cpp")], check=True, env=env)
    return str(path)


@pytest.mark.parametrize("threads", [None, "1", "4"])
@pytest.mark.parametrize("nt", ["1", "0"])
def test_copy_pool(exe, nt, threads):
    env = dict(os.environ, INTERPN_B200_COPY_NT=nt)
    if threads is not None:
        env["INTERPN_B200_COPY_THREADS"] = threads
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
