"""Drive the UNMODIFIED reference as a live differential oracle.

TEST INFRASTRUCTURE ONLY.  Works only where ``/root/reference`` is mounted
(the build container); nothing that runs on the GPU box may import this
module.  It imports the reference's own modules in place -- no reference source
is copied -- and feeds them pre-computed logits rows through a stub model, so
the language model is out of the comparison.

Shims (each one works around a defect of the reference in this image, none
changes the coder arithmetic):
  1. ``bitarray`` is not installed; only ``expansion_ratio`` uses it
     (code_base/utils.py:135-140) -> empty stub module.
  2. ``code_base`` must be first on ``sys.path`` so ``from utils import ...``
     resolves to code_base/utils.py, not the root shim.
  3. ``decode_arithmetic`` reads ``max_positions`` (code_base/arithmetic.py:257)
     which only ``encode_arithmetic`` defines (:91-94) -> module attribute.
  4. stub model returning ``.logits`` and ``past_key_values=None`` so the
     DynamicCache helpers (:12-41, broken on transformers 5.5) are never entered.
  5. stub tokenizer (space-joined ints) keeps the BPE-repair branch inert.
"""

from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Callable, List, Sequence

import numpy as np

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "code_base"))


_MODS = {}


def _load_code_base():
    if "arithmetic" in _MODS:
        return _MODS
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    sys.modules.setdefault("bitarray", types.ModuleType("bitarray"))            # shim 1
    cb = os.path.join(REFERENCE_ROOT, "code_base")
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in ("utils", "arithmetic", "huffman",
                                                  "huffman_baseline", "block_baseline")
                  if k in sys.modules}
    try:
        sys.path.insert(0, cb)                                                  # shim 2
        for name in ("utils", "arithmetic", "huffman", "huffman_baseline", "block_baseline"):
            _MODS[name] = importlib.import_module(name)
    finally:
        sys.path[:] = saved_path
        for name in ("utils", "arithmetic", "huffman", "huffman_baseline", "block_baseline"):
            sys.modules.pop(name, None)
        sys.modules.update(saved_mods)
    _MODS["arithmetic"].max_positions = 1024                                    # shim 3
    return _MODS


class _Out:
    def __init__(self, logits):
        self.logits = logits
        self.past_key_values = None


class StubModel:                                                                # shim 4
    """Returns ``rows(t)`` as the logits of call number ``t``."""

    config = types.SimpleNamespace(n_positions=1024)

    def __init__(self, rows: Callable[[int], np.ndarray], legacy_tuple: bool = False):
        self.rows, self.t, self.legacy = rows, 0, legacy_tuple

    def __call__(self, ids, past_key_values=None, use_cache=True, position_ids=None, past=None):
        import torch

        row = torch.from_numpy(np.array(self.rows(self.t), dtype=np.float32, copy=True))
        self.t += 1
        logits = row.view(1, 1, -1).repeat(1, ids.shape[1], 1)
        if self.legacy:                       # huffman/bins: model(x, past=past) -> (logits, past)
            return logits, []
        return _Out(logits)


class StubTokenizer:                                                            # shim 5
    def __init__(self, vocab: int = 0):
        self.decoder = {i: " %d" % i for i in range(vocab)}

    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)

    def encode(self, text):
        return [int(t) for t in text.split()]


CONTEXT = [1, 2, 3]


def ref_encode_arithmetic(rows, message: Sequence[int], *, temp, precision, topk):
    """Live ``code_base.encode_arithmetic`` (code_base/arithmetic.py:78)."""
    A = _load_code_base()["arithmetic"]
    toks, nll, kl, wpb, hq = A.encode_arithmetic(StubModel(rows), StubTokenizer(), list(message), CONTEXT,
                                                 device="cpu", temp=temp, precision=precision, topk=topk)
    return toks, (nll, kl, wpb, hq)


def ref_decode_arithmetic(rows, tokens: Sequence[int], *, temp, precision, topk) -> List[int]:
    """Live ``code_base.decode_arithmetic`` (code_base/arithmetic.py:220)."""
    A = _load_code_base()["arithmetic"]
    tok = StubTokenizer()
    return A.decode_arithmetic(StubModel(rows), tok, tok.decode(tokens), CONTEXT,
                               device="cpu", temp=temp, precision=precision, topk=topk)


def ref_encode_huffman(rows, message, bits_per_word: int):
    """Live ``encode_huffman`` (code_base/huffman_baseline.py:7)."""
    H = _load_code_base()["huffman_baseline"]
    out = H.encode_huffman(StubModel(rows, legacy_tuple=True), StubTokenizer(), list(message), CONTEXT,
                           bits_per_word, device="cpu")
    return out[0], out[1:]


def ref_decode_huffman(rows, tokens, bits_per_word: int, vocab: int):
    """Live ``decode_huffman`` (code_base/huffman_baseline.py:73)."""
    H = _load_code_base()["huffman_baseline"]
    tok = StubTokenizer(vocab)
    return H.decode_huffman(StubModel(rows, legacy_tuple=True), tok, tok.decode(tokens), CONTEXT,
                            bits_per_word, device="cpu")


def ref_get_bins(vocab: int, block_size: int):
    """Live ``get_bins`` (code_base/block_baseline.py:9)."""
    return _load_code_base()["block_baseline"].get_bins(vocab, block_size)


def ref_encode_block(rows, message, block_size: int, vocab: int):
    """Live ``encode_block`` (code_base/block_baseline.py:26)."""
    B = _load_code_base()["block_baseline"]
    b2w, w2b = B.get_bins(vocab, block_size)
    out = B.encode_block(StubModel(rows, legacy_tuple=True), StubTokenizer(), list(message), CONTEXT,
                         block_size, b2w, w2b, device="cpu")
    return out[0], out[1:]


def ref_decode_block(rows, tokens, block_size: int, vocab: int):
    """Live ``decode_block`` (code_base/block_baseline.py:99)."""
    B = _load_code_base()["block_baseline"]
    b2w, w2b = B.get_bins(vocab, block_size)
    tok = StubTokenizer(vocab)
    return B.decode_block(StubModel(rows, legacy_tuple=True), tok, tok.decode(tokens), CONTEXT,
                          block_size, b2w, w2b, device="cpu")


# ---------------------------------------------------------------------------
# (B) src/neuralstego rank codec
# ---------------------------------------------------------------------------
def _load_src():
    if "codec_arith" in _MODS:
        return _MODS
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.append(src)
    _MODS["codec_arith"] = importlib.import_module("neuralstego.codec.arithmetic")
    return _MODS


class RowsLM:
    """``next_token_probs`` provider: fp64 softmax of ``rows(len(context) - base)``.

    Mirrors ``_ModelAdapter.next_token_probs`` (src/neuralstego/lm/arithmetic.py:69-73)
    with the model replaced by the row source.
    """

    def __init__(self, rows, base: int, temperature: float = 1.0):
        self.rows, self.base, self.temperature = rows, base, temperature

    def next_token_probs(self, context_ids):
        import torch

        t = len(tuple(context_ids)) - self.base
        logits = torch.from_numpy(np.array(self.rows(t), dtype=np.float32)).to(torch.float64)
        logits = logits / self.temperature
        return torch.nn.functional.softmax(logits, dim=-1).numpy()


def ref_rank_encode(rows, payload: bytes, *, temperature=1.0, quality=None):
    """Live ``encode_with_lm`` (src/neuralstego/codec/arithmetic.py:122)."""
    C = _load_src()["codec_arith"]
    state = {}
    toks = C.encode_with_lm(payload, RowsLM(rows, len(CONTEXT), temperature), context=CONTEXT,
                            quality=quality, state=state)
    return toks, state


def ref_rank_decode(rows, tokens, state, *, temperature=1.0, quality=None) -> bytes:
    """Live ``decode_with_lm`` (src/neuralstego/codec/arithmetic.py:172)."""
    C = _load_src()["codec_arith"]
    return C.decode_with_lm(list(tokens), RowsLM(rows, len(CONTEXT), temperature), context=CONTEXT,
                            quality=quality, state=dict(state))
