"""CPU tier: the C-ABI shared library loads without a GPU and exports every entry point include/swinvox_b200.h
declares; the ctypes mirrors agree with the C struct layouts; a compute call without a device fails loudly."""
import ctypes as C
import os
import re

import pytest

from swinvox_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "swinvox_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:int|void|const char\*|svx_plan\*)\s+(svx_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    assert len(names) >= 39 and "svx_gemm" in names and "svx_plan_run" in names and "svx_binvox_decode" in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail(f"{_lib.LIB_PATH} missing: run `python __graft_entry__.py build`")
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_mirror_matches_struct_layouts_and_names():
    lib = _lib.bind(_lib.LIB_PATH)     # raises on any size / symbol / ABI-version mismatch
    assert lib.svx_abi_version() == 2
    declared = set(declared_functions())
    bound = set(_lib.ALL_SYMBOLS) | set(_lib.IO_SYMBOLS)
    assert bound <= declared, f"bound in python but not declared in the header: {sorted(bound - declared)}"
    assert declared <= bound, f"declared but never bound: {sorted(declared - bound)}"


def test_null_descriptor_is_an_error_not_a_crash():
    lib = _lib.bind(_lib.LIB_PATH)
    assert lib.svx_gemm(None, None) != 0
    assert b"null descriptor" in lib.svx_last_error()
    assert lib.svx_binvox_decode(None, None) != 0


def test_mlp_rejects_unsupported_shapes_before_touching_the_device():
    """svx_mlp validates its descriptor first (widths, hidden = 4C, alignment, LayerNorm only for C = 96): a bad
    description is a return code + message, never a launch."""
    lib = _lib.bind(_lib.LIB_PATH)
    assert lib.svx_mlp(None, None) != 0 and b"null descriptor" in lib.svx_last_error()
    buf = (C.c_float * 64)()
    base = C.addressof(buf)
    base += (16 - base % 16) % 16

    def desc(Cc, hidden, ln=False):
        d = _lib.MlpDesc()
        d.x = d.W1 = d.b1 = d.W2 = d.b2 = d.residual = d.out = base
        d.ldx = d.ldo = Cc
        d.M, d.C, d.hidden = 128, Cc, hidden
        if ln:
            d.ln_gamma = d.ln_beta = base
            d.ln_eps = 1e-5
        return d

    for bad in (desc(128, 512), desc(96, 192), desc(384, 1536), desc(192, 768, ln=True)):
        plan = lib.svx_plan_create()
        try:
            assert lib.svx_plan_add_mlp(plan, C.byref(bad)) != 0
            assert b"mlp" in lib.svx_last_error()
            assert lib.svx_plan_num_ops(plan) == 0
        finally:
            lib.svx_plan_destroy(plan)
