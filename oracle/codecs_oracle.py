"""Numpy/Python restatement of the comparison codecs.  TEST INFRASTRUCTURE ONLY.

* rank codec (B): src/neuralstego/codec/arithmetic.py:122-231, :370-385
* Huffman baseline: code_base/huffman_baseline.py:7-71, :73-165; code_base/huffman.py:12-76
* bins baseline: code_base/block_baseline.py:9-24, :26-97, :99-189

All take ``rows(t)`` = the fp32 logits row of step ``t`` (see ac_oracle.py).
Tie-break everywhere: equal logits -> lower token id first (the reference's
sorts are unstable there).
"""

from __future__ import annotations

import heapq
import math
import os
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

from .ac_oracle import DEC_MASK, bits2int, int2bits, mask_row, sort_desc


# ----------------------------------------------------------------------------
# (B) rank codec
# ----------------------------------------------------------------------------
def bytes_to_bits_msb(payload: bytes) -> List[int]:
    """codec/arithmetic.py:553-558."""
    return [(byte >> shift) & 1 for byte in payload for shift in range(7, -1, -1)]


def bits_to_bytes_msb(bits: Sequence[int]) -> bytes:
    """BitWriter.to_bytes, codec/arithmetic.py:99-119."""
    bits = list(bits)
    bits += [0] * ((-len(bits)) % 8)
    out = bytearray()
    for i in range(0, len(bits), 8):
        b = 0
        for bit in bits[i:i + 8]:
            b = (b << 1) | int(bit)
        out.append(b)
    return bytes(out)


def rank_distribution(row: np.ndarray, temperature: float, top_k=None, top_p=None, min_prob=None,
                      cap_per_token_bits=None) -> np.ndarray:
    """fp64 softmax of logits/temperature (lm/arithmetic.py:69-73) + the quality filters of
    codec/quality.py:57-105 (top_k, top_p, min_prob: each a prefix of the descending order, evaluated on
    the unfiltered probabilities, then one renormalisation) + the entropy cap of :108-141."""
    x = row.astype(np.float64) / float(temperature)
    e = np.exp(x - x.max())
    p = e / e.sum()
    if top_k is not None or top_p is not None or min_prob is not None:
        order = np.argsort(-p, kind="stable")
        keep = np.ones(p.size, dtype=bool)
        if top_k is not None:                                           # :76-81
            keep[:] = False
            keep[order[: min(int(top_k), p.size)]] = True
        if top_p is not None:                                           # :85-91
            cutoff = int(np.searchsorted(np.cumsum(p[order]), top_p, side="left"))
            inside = np.zeros(p.size, dtype=bool)
            inside[order[: cutoff + 1]] = True
            keep &= inside
        if min_prob is not None:                                        # :93-96
            keep &= p >= min_prob
        if not keep.any():
            raise ValueError("Quality policies removed all probability mass")   # :98-99
        f = np.where(keep, p, 0.0)
        p = f / f.sum()
    if cap_per_token_bits is not None:
        p = cap_bits(p, int(cap_per_token_bits))
    return p


def _entropy_bits(p: np.ndarray) -> float:
    v = p[p > 0.0]
    return float(-(v * np.log2(v)).sum()) if v.size else 0.0


def cap_bits(p: np.ndarray, cap: int) -> np.ndarray:
    """cap_bits_per_token (quality.py:108-141): 60-step bisection on a sharpening temperature.  For the rank
    codec only the support and the order of the result matter; both are unchanged unless the sharpening
    underflows the tail or resurrects filtered tokens through the ``+1e-12`` (then the reference's order among
    the equal probabilities is its unstable argsort's -- unpinned)."""
    p = p / p.sum()
    if _entropy_bits(p) <= cap:
        return p
    low, high, target = 1e-6, 1.0, p
    for _ in range(60):
        mid = (low + high) / 2.0
        z = np.log(p + 1e-12) / mid
        z -= z.max()
        c = np.exp(z)
        c = c / c.sum()
        if _entropy_bits(c) > cap:
            high = mid
        else:
            target, low = c, mid
    return target


def rank_tokens(p: np.ndarray) -> Tuple[np.ndarray, int]:
    """codec/arithmetic.py:370-385: tokens with p > 0 by descending p, top 2^capacity."""
    ids = np.nonzero(p > 0)[0]
    order = ids[np.argsort(-p[ids], kind="stable")]
    capacity = int(math.floor(math.log2(order.size)))
    if capacity <= 0:
        return order, 0
    return order[: 1 << capacity], capacity


def rank_encode(rows: Callable[[int], np.ndarray], payload: bytes, *, temperature: float = 1.0, top_k=None, **quality):
    """encode_with_lm (codec/arithmetic.py:122-169). Returns (tokens, history, total_bits)."""
    bits = bytes_to_bits_msb(payload)
    total = len(bits)
    pos, t = 0, 0
    tokens: List[int] = []
    history: List[int] = []
    while pos < total:                                                  # :146
        ranked, capacity = rank_tokens(rank_distribution(rows(t), temperature, top_k, **quality))
        if capacity <= 0:
            raise ValueError("no capacity")
        chunk = bits[pos: pos + capacity]
        consumed = len(chunk)
        chunk = chunk + [0] * (capacity - consumed)                     # BitReader zero padding :73-76
        index = 0
        for b in chunk:
            index = (index << 1) | b
        tokens.append(int(ranked[index]))                               # :160
        history.append(consumed)
        pos += consumed
        t += 1
    return tokens, history, total


def rank_decode(rows, tokens: Sequence[int], history: Sequence[int], total_bits: int, *,
                temperature: float = 1.0, top_k=None, **quality) -> bytes:
    """decode_with_lm (codec/arithmetic.py:172-231)."""
    out: List[int] = []
    for t, tok in enumerate(tokens):
        ranked, capacity = rank_tokens(rank_distribution(rows(t), temperature, top_k, **quality))
        index = int(np.nonzero(ranked == tok)[0][0])                    # :211
        emitted = [(index >> s) & 1 for s in reversed(range(capacity))]  # :529-530
        out += emitted[: history[t]]                                    # :216
    return bits_to_bytes_msb(out[:total_bits])


# ----------------------------------------------------------------------------
# Huffman baseline
# ----------------------------------------------------------------------------
class _Node:
    """code_base/huffman.py:12-28 -- ordered by frequency only."""

    __slots__ = ("token", "freq", "left", "right")

    def __init__(self, token, freq):
        self.token, self.freq, self.left, self.right = token, freq, None, None

    def __lt__(self, other):
        return self.freq < other.freq


def huffman_tree(freqs: np.ndarray):
    """make_heap_from_array + merge_nodes + make_codes (huffman.py:43-76)."""
    heap: List[_Node] = []
    for idx in range(len(freqs)):
        heapq.heappush(heap, _Node(idx, freqs[idx]))
    while len(heap) > 1:
        n1 = heapq.heappop(heap)
        n2 = heapq.heappop(heap)
        merged = _Node(None, n1.freq + n2.freq)
        merged.left, merged.right = n1, n2
        heapq.heappush(heap, merged)
    root = heapq.heappop(heap)
    codes: Dict[int, str] = {}

    def walk(node, code):
        if node.token is not None:
            codes[node.token] = code
            return
        walk(node.left, code + "0")
        walk(node.right, code + "1")

    walk(root, "")
    return root, codes


def huffman_probs(row: np.ndarray, bits_per_word: int):
    """Top 2^b tokens and their fp32 probabilities (huffman_baseline.py:26-34)."""
    s, order = sort_desc(mask_row(row, DEC_MASK))
    z = s - s.max()
    logp = z - np.log(np.exp(z).sum(dtype=np.float32), dtype=np.float32)
    n = 1 << bits_per_word
    return order[:n], np.exp(logp[:n].astype(np.float32))


def huffman_encode(rows, message: Sequence[int], bits_per_word: int):
    """encode_huffman without finish_sent (huffman_baseline.py:7-71). Returns (tokens, bits_consumed)."""
    message = [int(b) for b in message]
    length = len(message)
    i, t = 0, 0
    tokens: List[int] = []
    while i < length:
        ids, probs = huffman_probs(rows(t), bits_per_word)
        node, _ = huffman_tree(probs)
        while node.token is None:                                       # :47-52
            if i >= length or message[i] == 0:
                node = node.left
            else:
                node = node.right
            i += 1
        tokens.append(int(ids[node.token]))
        t += 1
    return tokens, i


def huffman_decode(rows, tokens: Sequence[int], bits_per_word: int) -> List[int]:
    """decode_huffman (huffman_baseline.py:73-165) for in-range tokens."""
    out: List[int] = []
    for t, tok in enumerate(tokens):
        ids, probs = huffman_probs(rows(t), bits_per_word)
        hit = np.nonzero(ids == tok)[0]
        rank = int(hit[0]) if hit.size else 0                           # :149
        _, codes = huffman_tree(probs)
        out += [int(c) for c in codes[rank]]                            # :159
    return out


# ----------------------------------------------------------------------------
# bins baseline
# ----------------------------------------------------------------------------
def get_bins(vocab_size: int, block_size: int):
    """block_baseline.py:9-24 (numpy legacy RNG seeded with the block size)."""
    num_bins = 2 ** block_size
    words_per_bin = vocab_size / num_bins
    ordering = np.arange(vocab_size)
    np.random.seed(block_size)
    np.random.shuffle(ordering)
    bin2words = [ordering[int(i * words_per_bin): int((i + 1) * words_per_bin)] for i in range(num_bins)]
    word2bin = np.full(vocab_size, -1, dtype=np.int32)
    for j, words in enumerate(bin2words):
        word2bin[words] = j
    return bin2words, word2bin


def bins_encode(rows, message: Sequence[int], block_size: int, vocab: int):
    """encode_block without finish_sent (block_baseline.py:26-97)."""
    message = [int(b) for b in message]
    bin2words, _ = get_bins(vocab, block_size)
    i, t = 0, 0
    tokens: List[int] = []
    while i < len(message):
        row = mask_row(rows(t), DEC_MASK)
        words = np.sort(bin2words[bits2int(message[i: i + block_size])])  # :79
        tokens.append(int(words[np.argmax(row[words])]))                # :80-81 (lowest id among ties)
        i += block_size
        t += 1
    return tokens, i


def bins_decode(tokens: Sequence[int], block_size: int, vocab: int) -> List[int]:
    """decode_block (block_baseline.py:99-189) for tokens that are their bin's argmax."""
    _, word2bin = get_bins(vocab, block_size)
    out: List[int] = []
    for tok in tokens:
        out += int2bits(int(word2bin[tok]), block_size)                 # :183
    return out


# ----------------------------------------------------------------------------
# golden generation (called from make_golden.py, build container only)
# ----------------------------------------------------------------------------
CODEC_CASES = [
    dict(name="huffman_v2048_b3", kind="huffman", V=2048, T=24, scale=3.0, param=3, streams=4, bits=96),
    dict(name="huffman_v50257_b3", kind="huffman", V=50257, T=8, scale=3.0, param=3, streams=2, bits=64),
    dict(name="huffman_v2048_b5", kind="huffman", V=2048, T=24, scale=1.0, param=5, streams=3, bits=120),
    dict(name="bins_v2048_b3", kind="bins", V=2048, T=24, scale=3.0, param=3, streams=4, bits=96),
    dict(name="bins_v50257_b3", kind="bins", V=50257, T=8, scale=3.0, param=3, streams=2, bits=60),
    dict(name="bins_v50257_b5", kind="bins", V=50257, T=8, scale=3.0, param=5, streams=2, bits=60),
    dict(name="rank_v2048_t10", kind="rank", V=2048, T=24, scale=3.0, param=0, temperature=1.0, streams=4, bits=160),
    dict(name="rank_v42001_t08", kind="rank", V=42001, T=8, scale=2.5, param=0, temperature=0.8, streams=2, bits=240),
    dict(name="rank_v2048_topk64", kind="rank", V=2048, T=24, scale=3.0, param=64, temperature=1.0, streams=3, bits=160),
    dict(name="rank_v2048_topp90", kind="rank", V=2048, T=24, scale=3.0, param=0, temperature=1.0, streams=3, bits=160,
         top_p=0.9),
    dict(name="rank_v2048_minp", kind="rank", V=2048, T=24, scale=3.0, param=0, temperature=1.0, streams=3, bits=160,
         min_prob=2e-4),
    dict(name="rank_v42001_mix", kind="rank", V=42001, T=8, scale=2.5, param=6000, temperature=0.8, streams=2, bits=200,
         top_p=0.97, min_prob=1e-6),
    dict(name="rank_v2048_cap3", kind="rank", V=2048, T=24, scale=3.0, param=0, temperature=1.0, streams=2, bits=160,
         cap_per_token_bits=3),
]


def rank_quality(cfg) -> dict:
    """The quality keys of a rank case besides top_k (golden metadata -> oracle / kernel arguments)."""
    return {k: cfg[k] for k in ("top_p", "min_prob", "cap_per_token_bits") if cfg.get(k) is not None}


def make_codec_goldens(out_dir, logits_pool, message_bits, rows_for):
    import zlib

    from . import ref_harness as H

    metas = []
    for idx, cfg in enumerate(CODEC_CASES):
        seed = 2000 + idx
        pool = logits_pool(seed, cfg["T"], cfg["V"], cfg["scale"])
        data = dict(pool_seed=seed, pool_crc=zlib.crc32(pool.tobytes()) & 0xFFFFFFFF)
        for s in range(cfg["streams"]):
            rows = rows_for(pool, s)
            msg = message_bits(seed * 1000 + s, cfg["bits"] - (8 * (s % 2) if cfg["kind"] == "rank" else s % 3))
            if cfg["kind"] == "huffman":
                ref_tok, _ = H.ref_encode_huffman(rows, msg.tolist(), cfg["param"])
                ref_bits = H.ref_decode_huffman(rows, ref_tok, cfg["param"], cfg["V"])
                tok, used = huffman_encode(rows, msg.tolist(), cfg["param"])
                bits = huffman_decode(rows, ref_tok, cfg["param"])
                assert tok == ref_tok, (cfg["name"], s)
                assert bits == list(ref_bits), (cfg["name"], s)
                assert bits[: len(msg)] == msg.tolist()
            elif cfg["kind"] == "bins":
                ref_tok, _ = H.ref_encode_block(rows, msg.tolist(), cfg["param"], cfg["V"])
                ref_bits = H.ref_decode_block(rows, ref_tok, cfg["param"], cfg["V"])
                tok, used = bins_encode(rows, msg.tolist(), cfg["param"], cfg["V"])
                bits = bins_decode(ref_tok, cfg["param"], cfg["V"])
                assert tok == ref_tok, (cfg["name"], s)
                assert bits == list(ref_bits), (cfg["name"], s)
            else:
                payload = bits_to_bytes_msb(msg.tolist())
                top_k = cfg["param"] or None
                extra = rank_quality(cfg)
                quality = dict(extra, **({"top_k": top_k} if top_k else {})) or None
                ref_tok, state = H.ref_rank_encode(rows, payload, temperature=cfg["temperature"], quality=quality)
                ref_payload = H.ref_rank_decode(rows, ref_tok, state, temperature=cfg["temperature"], quality=quality)
                tok, hist, total = rank_encode(rows, payload, temperature=cfg["temperature"], top_k=top_k, **extra)
                back = rank_decode(rows, ref_tok, hist, total, temperature=cfg["temperature"], top_k=top_k, **extra)
                assert tok == list(ref_tok), (cfg["name"], s)
                if "cap_per_token_bits" in extra:      # the cap leaves order and support alone: same tokens without it
                    assert tok == rank_encode(rows, payload, temperature=cfg["temperature"], top_k=top_k)[0]
                assert tuple(hist) == tuple(state["history"]), (cfg["name"], s)
                assert back == ref_payload == payload, (cfg["name"], s)
                data["history_%d" % s] = np.asarray(hist, dtype=np.int32)
                ref_bits = bytes_to_bits_msb(ref_payload)
            data["msg_%d" % s] = msg
            data["tokens_%d" % s] = np.asarray(ref_tok, dtype=np.int32)
            data["decoded_%d" % s] = np.asarray(list(ref_bits), dtype=np.uint8)
        np.savez_compressed(os.path.join(out_dir, cfg["name"] + ".npz"), **data)
        metas.append(cfg)
        print(cfg["kind"], cfg["name"], "ok")
    return {"codecs": metas}
