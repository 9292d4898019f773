"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/ns_coder.h declares.
No compute call is made here (no GPU in the CPU suite)."""
import ctypes as C
import os
import re

import pytest

from neuralsteganography_b200 import _native as N
from neuralsteganography_b200.build import build_native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_native()
    return N.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ns_coder.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ns_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert "ns_ac_encode_step" in names and "ns_ac_decode_step" in names
    for name in names:
        assert hasattr(lib, name), "missing export: " + name


def test_version_and_capacity(lib):
    assert lib.ns_version() == 1
    assert lib.ns_ac_max_vocab() >= 50257        # GPT-2 vocabulary must fit one CTA's shared memory


def test_struct_layout_matches_header(lib):
    # sizeof(ns_ac_params) as the compiler sees it, via a probe symbol
    assert C.sizeof(N.AcParams) == lib.ns_sizeof_ac_params()
    if hasattr(lib, "ns_sizeof_codec_params"):
        assert C.sizeof(N.CodecParams) == lib.ns_sizeof_codec_params()


def test_argument_validation_without_gpu(lib):
    p = N.AcParams()
    assert lib.ns_ac_encode_step(C.byref(p), None) == -1          # NS_E_NULL
    assert b"NULL" in lib.ns_last_error_string()
    p.logits = 16; p.lo = 16; p.hi = 16; p.B = 1; p.V = 10 ** 6; p.ld = 10 ** 6
    assert lib.ns_ac_encode_step(C.byref(p), None) == -3          # NS_E_VOCAB
    p.V = 1000; p.ld = 1000; p.precision = 99
    assert lib.ns_ac_encode_step(C.byref(p), None) == -2          # NS_E_RANGE


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(N.NativeLibraryError):
        N.load(str(tmp_path / "nope.so"))
