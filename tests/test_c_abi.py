"""The boundary is a real C ABI: a plain-C host program (no Python, no torch) links libfs2b200.so through
include/fs2_b200.h.  CPU: it must compile and link.  GPU: it must run the LengthRegulator bit-exactly."""
import os
import shutil
import subprocess

import pytest

from conftest import REPO

SRC = os.path.join(REPO, "tests", "c_abi", "length_regulator_host.c")
LIBDIR = os.path.join(REPO, "fastspeech2_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(tmp_path):
    exe = str(tmp_path / "length_regulator_host")
    cmd = ["gcc", "-std=c99", "-Wall", "-I", os.path.join(REPO, "include"), "-I", os.path.join(CUDA, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lfs2b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart", f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{CUDA}/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_host_program_compiles_and_links(tmp_path):
    build(tmp_path)


@pytest.mark.gpu
def test_c_host_program_runs_bit_exact(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C_ABI_OK" in r.stdout, r.stdout + r.stderr
    assert "sm_100a" in r.stdout
