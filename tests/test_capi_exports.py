"""The C-ABI library builds for sm_100a (nvcc cross-compiles on the CPU box), loads, and exports
every entry point declared in include/mapdn_b200.h. No compute calls - no GPU needed."""
import ctypes
import os
import re

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mapdn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mapdn_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    fns = declared_functions()
    for must in ("mapdn_create", "mapdn_destroy", "mapdn_reset", "mapdn_step", "mapdn_step_host", "mapdn_get_obs",
                 "mapdn_get_state", "mapdn_get_field", "mapdn_solve"):
        assert must in fns


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    for fn in declared_functions():
        assert hasattr(L, fn), f"{fn} declared in the header but not exported"
    L.mapdn_abi_version.restype = ctypes.c_int32
    assert L.mapdn_abi_version() == 1


def test_binding_covers_the_header(built_lib):
    from mapdn_b200 import _capi
    assert sorted(_capi.EXPORTS) == declared_functions()
    assert _capi.lib().mapdn_abi_version() == 1


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors == sizeof computed by the C compiler for the header structs."""
    import subprocess, tempfile
    from mapdn_b200 import _capi
    prog = r'''
#include <stdio.h>
#include "mapdn_b200.h"
int main(void){ printf("%zu %zu %zu %zu\n", sizeof(mapdn_net_desc), sizeof(mapdn_profile_desc), sizeof(mapdn_cfg), sizeof(mapdn_dims)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_capi.NetDescC), ctypes.sizeof(_capi.ProfileDescC), ctypes.sizeof(_capi.CfgC),
                     ctypes.sizeof(_capi.DimsC)]


def test_no_oracle_import_in_product():
    """The product never routes through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "mapdn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
