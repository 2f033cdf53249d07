"""The compile-time experiments (tuning knobs) must keep compiling. Opt-in (MAPDN_TEST_VARIANTS=1): each
variant is a full nvcc build of every kernel instantiation, about a minute apiece."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

MACROS = ["MAPDN_EXP_CHILD_LOADS_FIRST", "MAPDN_EXP_HELPER_V2", "MAPDN_PROFILE"]


@pytest.mark.skipif(os.environ.get("MAPDN_TEST_VARIANTS") != "1", reason="opt-in: MAPDN_TEST_VARIANTS=1")
@pytest.mark.parametrize("macro", MACROS)
def test_variant_compiles(macro, tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    out = tmp_path / "variant.so"
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--shared", "-Xcompiler",
                        "-fPIC", "--expt-relaxed-constexpr", f"-D{macro}", "-o", str(out),
                        os.path.join(ROOT, "mapdn_b200", "csrc", "mapdn_b200.cu")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.stat().st_size > 0
