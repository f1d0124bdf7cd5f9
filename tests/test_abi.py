"""CPU: libgrmkm.so loads and exports every symbol include/grmkm.h declares; no compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from grm_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "grmkm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(grmkm_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(native.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = native.load()
    for name in header_symbols():
        assert getattr(lib, name) is not None, name
    assert lib.grmkm_abi_version() == native.ABI_VERSION
    m = re.search(r"#define GRMKM_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "grmkm.h")).read())
    assert int(m.group(1)) == native.ABI_VERSION


def test_struct_layouts_match_the_header():
    # grmkm_config: 8 x u32/i32 + pointer; grmkm_stats: 6 x u64 + 4 x u32 + 4 x u64 + 3 x u64 + 2 x u32 + u64;
    # grmkm_times: 12 floats
    assert C.sizeof(native.Config) == 8 * 4 + 8
    assert C.sizeof(native.Stats) == 6 * 8 + 4 * 4 + 4 * 8 + 3 * 8 + 2 * 4 + 8
    assert C.sizeof(native.Times) == 12 * 4


def test_argument_errors_do_not_need_a_device():
    lib = native.load()
    ctx = C.c_void_p()
    bad = native.Config(C.sizeof(native.Config), 33, 1, 0, 0, -1, 0, 0, None)
    assert lib.grmkm_create(C.byref(bad), C.byref(ctx)) == native.E_UNSUPPORTED_K
    assert b"1..32" in lib.grmkm_last_error(None)
    assert lib.grmkm_create(None, C.byref(ctx)) == native.E_INVALID
    assert lib.grmkm_build(None) == native.E_INVALID


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product path fails loudly (there is no CPU code path to fall back to)."""
    lib = native.load()
    if lib.grmkm_device_count() > 0:
        pytest.skip("a GPU is visible here")
    from grm_b200.builder import KmerMatrixBuilder
    with pytest.raises(native.GrmkmError) as e:
        KmerMatrixBuilder(k=31)
    assert e.value.code == native.E_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "genomic-resistance-mapping-grm-_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "grm_oracle" not in text and "libgrmoracle" not in text, f
