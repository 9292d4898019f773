"""ctypes binding of ``libns_coder.so`` (C ABI declared in ``include/ns_coder.h``).

The library is built in-tree by ``__graft_entry__.build()`` /
``neuralsteganography_b200.build.build_native()``.  There is no CPU fallback:
if the library or a CUDA device is missing every call raises
:class:`NativeLibraryError`.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NS_CODER_LIB") or os.path.join(_PKG_DIR, "libns_coder.so")   # override: kernel experiments

NS_OK = 0
PHASE_CODING, PHASE_TAIL, PHASE_DONE = 0, 1, 2
ST_OUT_OF_RANGE, ST_BIN_OVERFLOW, ST_EST_RETRY, ST_TOKEN_OVERFLOW, ST_RANK_DEFER = 1, 2, 4, 8, 16

EXPORTS = (
    "ns_version", "ns_last_error_string", "ns_ac_max_vocab",
    "ns_ac_encode_step", "ns_ac_decode_step", "ns_ac_debug_bins",
    "ns_rank_encode_step", "ns_rank_decode_step",
    "ns_huffman_encode_step", "ns_huffman_decode_step",
    "ns_bins_encode_step", "ns_bins_decode_step",
)


class NativeLibraryError(RuntimeError):
    """The CUDA extension is missing, failed to load, or reported an error."""


class AcParams(C.Structure):
    """Mirror of ``ns_ac_params`` (include/ns_coder.h)."""

    _fields_ = [
        ("logits", C.c_void_p), ("ld", C.c_int64), ("B", C.c_int32), ("V", C.c_int32),
        ("temp", C.c_double), ("precision", C.c_int32), ("topk", C.c_int32),
        ("mask_id", C.c_int32 * 2),
        ("lo", C.c_void_p), ("hi", C.c_void_p), ("phase", C.c_void_p), ("status", C.c_void_p),
        ("ntok", C.c_void_p), ("token_cap", C.c_int32), ("ntok_total", C.c_void_p),
        ("msg", C.c_void_p), ("msg_stride", C.c_int64), ("msg_len", C.c_void_p),
        ("cursor", C.c_void_p), ("token_out", C.c_void_p), ("token_stride", C.c_int64),
        ("finish_sent", C.c_int32), ("sent_end", C.c_void_p),
        ("token_in", C.c_void_p), ("is_last", C.c_void_p),
        ("out_bits", C.c_void_p), ("out_stride", C.c_int64), ("out_len", C.c_void_p),
        ("nbits_out", C.c_void_p), ("trace", C.c_void_p),
        ("slow_ws", C.c_void_p), ("force_exact", C.c_int32), ("prof", C.c_void_p), ("stats", C.c_void_p),
        ("rank_ws", C.c_void_p), ("scratch_stride", C.c_int64), ("scratch_slots", C.c_int32), ("variant", C.c_int32),
    ]


class CodecParams(C.Structure):
    """Mirror of ``ns_codec_params`` (include/ns_coder.h): rank / Huffman / bins codecs."""

    _fields_ = [
        ("logits", C.c_void_p), ("ld", C.c_int64), ("B", C.c_int32), ("V", C.c_int32),
        ("temp", C.c_double), ("param", C.c_int32), ("topk", C.c_int32),
        ("mask_id", C.c_int32 * 2),
        ("phase", C.c_void_p), ("status", C.c_void_p),
        ("ntok", C.c_void_p), ("token_cap", C.c_int32), ("ntok_total", C.c_void_p),
        ("msg", C.c_void_p), ("msg_stride", C.c_int64), ("msg_len", C.c_void_p),
        ("cursor", C.c_void_p), ("token_out", C.c_void_p), ("token_stride", C.c_int64),
        ("token_in", C.c_void_p),
        ("out_bits", C.c_void_p), ("out_stride", C.c_int64), ("out_len", C.c_void_p),
        ("total_bits", C.c_void_p), ("nbits_out", C.c_void_p), ("lut", C.c_void_p),
        ("top_p", C.c_double), ("min_prob", C.c_double),
    ]


_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """Load the shared library (once) and declare its prototypes."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise NativeLibraryError(
            "CUDA extension %s not found; run `python -c 'import __graft_entry__ as g; g.build()'`" % p)
    try:
        lib = C.CDLL(p)
    except OSError as exc:
        raise NativeLibraryError("cannot load %s: %s" % (p, exc)) from exc
    lib.ns_version.restype = C.c_int
    lib.ns_last_error_string.restype = C.c_char_p
    lib.ns_ac_max_vocab.restype = C.c_int
    lib.ns_sizeof_ac_params.restype = C.c_int
    if hasattr(lib, "ns_sizeof_codec_params"):
        lib.ns_sizeof_codec_params.restype = C.c_int
    for name in ("ns_ac_encode_step", "ns_ac_decode_step"):
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(AcParams), C.c_void_p]
        fn.restype = C.c_int
    lib.ns_ac_debug_bins.argtypes = [C.POINTER(AcParams), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ns_ac_debug_bins.restype = C.c_int
    for name in ("ns_rank_encode_step", "ns_rank_decode_step", "ns_huffman_encode_step",
                 "ns_huffman_decode_step", "ns_bins_encode_step", "ns_bins_decode_step"):
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.argtypes = [C.POINTER(CodecParams), C.c_void_p]
            fn.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != NS_OK:
        msg = load().ns_last_error_string()
        raise NativeLibraryError("%s failed with code %d: %s" % (what, rc, (msg or b"").decode()))


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()
