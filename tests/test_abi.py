"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/gmlm_b200.h declares; the ctypes table covers exactly that set;
argument validation that needs no GPU returns the documented error codes."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "gmlm_b200.h"


def declared_symbols():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gmlm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("gmlm_degree_f32", "gmlm_edge_type_bucket", "gmlm_csr_build", "gmlm_csr_transpose",
                 "gmlm_spmm_csr", "gmlm_graphnorm_fwd", "gmlm_graphnorm_bwd_apply", "gmlm_soft_mask_fwd",
                 "gmlm_soft_mask_bwd"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_built):
    lib = C.CDLL(str(lib_built))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/gmlm_b200.h but not exported"


def test_ctypes_table_matches_header(lib_built):
    from gmlm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.gmlm_abi_version() >= 1


def test_library_is_sm100a_and_has_no_torch_dependency(lib_built):
    out = subprocess.run(["cuobjdump", "--list-elf", str(lib_built)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", str(lib_built)], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd


def test_argument_validation_without_gpu(lib_built):
    """These calls fail validation before touching the device."""
    from gmlm_b200 import _lib
    lib = _lib.load()
    null = C.c_void_p(0)
    # bad dtype code
    rc = lib.gmlm_spmm_csr(null, 7, 4, 4, null, null, null, 1, 1, 0, null, 0, 0, 0, 0, null, null, null, null, null, null, 4, null)
    assert rc == 1 and b"dtype" in lib.gmlm_last_error()
    # weighted mode without weights
    rc = lib.gmlm_spmm_csr(null, 0, 4, 4, null, null, null, 1, 1, 2, null, 0, 0, 0, 0, null, null, null, null, null, null, 4, null)
    assert rc == 1
    # too many bucket bounds
    rc = lib.gmlm_edge_type_bucket(null, 0, null, 0, (C.c_int32 * 9)(*range(9)), 9, null, null)
    assert rc == 1
    with pytest.raises(_lib.GmlmError):
        _lib.check(rc, "edge_type_bucket")
