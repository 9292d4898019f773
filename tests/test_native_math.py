"""Host build of csrc/ns_math.cuh (the scalar pieces the kernels run per element / per stream)
against the oracle.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from neuralsteganography_b200.build import build_hostmath
from oracle import ac_oracle as O


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(build_hostmath())
    lib.nsh_interval_update.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.nsh_read_bits.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int]
    lib.nsh_read_bits.restype = C.c_uint64
    lib.nsh_write_bits.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_int]
    lib.nsh_orderable.argtypes = [C.c_float]
    lib.nsh_orderable.restype = C.c_uint32
    return lib


def test_exp64_within_two_ulp_and_monotone(host):
    rng = np.random.default_rng(0)
    a = np.concatenate([-np.abs(rng.standard_normal(400_000)) * 8, -rng.uniform(0, 707.9, 200_000), [0.0, -707.99]])
    out = np.empty_like(a)
    host.nsh_exp64(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_int64(len(a)))
    ref = np.exp(a)
    err = np.abs(out - ref) / np.spacing(ref)
    assert err.max() <= 2.0                  # libm itself is <= 1 ulp: the two differ by at most 2
    s = np.sort(a)
    o = np.empty_like(s)
    host.nsh_exp64(s.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_int64(len(s)))
    assert (np.diff(o) >= 0).all()
    z = np.array([-708.5, -1e20, -np.inf, np.nan])
    oz = np.empty_like(z)
    host.nsh_exp64(z.ctypes.data_as(C.c_void_p), oz.ctypes.data_as(C.c_void_p), C.c_int64(len(z)))
    assert (oz == 0).all()                   # masked tokens get probability exactly 0


def test_interval_update_matches_oracle(host):
    rng = np.random.default_rng(1)
    for precision in (2, 8, 16, 26, 32, 40, 48):
        top = 1 << precision
        for _ in range(400):
            nb = int(rng.integers(0, top - 1))
            width = int(min(top - nb, max(1, int(rng.integers(1, top)) >> int(rng.integers(0, precision)))))
            nt = nb + width
            lo, hi = C.c_uint64(), C.c_uint64()
            n = host.nsh_interval_update(nb, nt, precision, C.byref(lo), C.byref(hi))
            en, elo, ehi, _, _ = O.interval_update(nb, nt, precision)
            assert (n, lo.value, hi.value) == (en, elo, ehi), (precision, nb, nt)


def test_bit_read_write_roundtrip(host):
    rng = np.random.default_rng(2)
    bits = rng.integers(0, 2, 500).tolist()
    words = np.zeros(20, dtype=np.uint32)
    pos = 0
    while pos < len(bits):
        cnt = int(min(len(bits) - pos, rng.integers(1, 49)))
        val = 0
        for b in bits[pos:pos + cnt]:
            val = (val << 1) | b
        host.nsh_write_bits(words.ctypes.data_as(C.c_void_p), pos, C.c_uint64(val), cnt)
        pos += cnt
    from neuralsteganography_b200.coder import pack_bits, unpack_bits
    packed, lens = pack_bits([bits])
    assert np.array_equal(packed[0, :16], words[:16])
    assert unpack_bits(packed, lens)[0] == bits
    for precision in (1, 7, 16, 26, 33, 48):
        for start in (0, 5, 31, 32, 60, 470, 499, 500, 520):
            got = host.nsh_read_bits(words.ctypes.data_as(C.c_void_p), start, len(bits), precision)
            window = bits[start:start + precision]
            window = window + [0] * (precision - len(window))
            assert got == O.bits2int(list(reversed(window)))   # code_base/arithmetic.py:168-171


def test_orderable_is_monotone(host):
    vals = np.array([-np.inf, -1e20, -3.5, -1e-30, 0.0, 1e-30, 2.0, 1e20], dtype=np.float32)
    keys = [host.nsh_orderable(float(v)) for v in vals]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
