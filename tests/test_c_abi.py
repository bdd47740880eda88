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


@pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.exists(os.path.join(CUDA, "bin", "cuobjdump")), reason="cuobjdump not available")
def test_library_is_sm100a_tcgen05_code():
    """What the shared library contains, read off its SASS (no GPU needed): sm_100a code only, 5th-generation tensor-core
    instructions (UTCHMMA = tcgen05.mma) fed by TMA (UTMALDG) with TMEM loads in the epilogues (LDTM), and not a single legacy
    warp-level HMMA -- i.e. the hot path is not a recompiled mma.sync / wmma kernel and not a library GEMM."""
    exe = shutil.which("cuobjdump") or os.path.join(CUDA, "bin", "cuobjdump")
    lib = os.path.join(LIBDIR, "libfs2b200.so")
    assert os.path.exists(lib), "build the library first (python -m fastspeech2_b200.build)"
    r = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    sass = r.stdout
    archs = {ln.split("=")[1].strip() for ln in sass.splitlines() if ln.startswith("arch =")}
    assert archs == {"sm_100a"}, archs
    count = {m: sass.count(m) for m in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "HMMA.")}
    assert count["UTCHMMA"] > 100 and count["UTMALDG"] > 100 and count["LDTM"] > 10 and count["UTCBAR"] > 10, count
    assert count["HMMA."] == 0, count
    for kernel in ("tap_gemm_tf32_kernel", "gemm_ln_cluster_kernel", "attention_f16_kernel", "length_gather_kernel", "row_norm_kernel"):
        assert kernel in sass, f"{kernel} missing from the library"
