"""B200-native steganographic coder step for NeuralSteganography (sm_100a).

Only the hot path lives here: the per-token arithmetic coder of
``code_base/arithmetic.py`` (and the rank / Huffman / bins comparison codecs)
as hand-written CUDA behind the C ABI of ``include/ns_coder.h``, plus the host
mirror of the reference's provider interface.  There is no CPU fallback.
"""

from ._native import NativeLibraryError  # noqa: F401

__all__ = ["NativeLibraryError"]
__version__ = "0.1.0"
