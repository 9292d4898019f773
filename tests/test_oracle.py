"""The oracle against the reference's known answers and the committed golden vectors.

The golden files were produced by the live, unmodified reference
(oracle/make_golden.py); these tests keep the numpy restatement pinned to them.
"""
import os
import zlib

import numpy as np
import pytest

from oracle import ac_oracle as O
from oracle import codecs_oracle as K
from oracle.inputs import logits_pool, message_bits, rows_for


def test_select_cutoff_k_known_answers():
    # reference: tests/codec/test_arithmetic_threshold.py:43-58
    probs = np.array([0.4, 0.35, 0.25])
    assert O.select_cutoff_k(probs, 0.1, 50) == 3
    assert O.select_cutoff_k(probs, 0.1, 2) == 2
    assert O.select_cutoff_k(np.array([0.9, 0.05, 0.05]), 0.1, 50) == 2   # max(2, 1)


def test_bit_helpers():
    assert O.bits2int([0, 1, 1, 1]) == 14                     # code_base/utils.py:42
    assert O.int2bits(14, 4) == [0, 1, 1, 1]
    assert O.int2bits(5, 0) == []
    # the reference's loop returns len-1 on identical lists
    assert O.num_same_from_beg([1, 0, 1], [1, 0, 1]) == 2
    assert O.num_same_from_beg([1, 0, 1], [1, 1, 1]) == 1
    assert O.num_same_from_beg([0, 0, 1], [1, 0, 1]) == 0


def test_interval_update_examples():
    n, lo, hi, bb, tb = O.interval_update(0b1010_0000, 0b1011_0000, 8)
    assert n == 4 and lo == 0 and hi == 256
    n, lo, hi, _, _ = O.interval_update(5, 6, 8)              # width-1 bin: capped at precision-1
    assert n == 7 and lo == 128 and hi == 256


def _load(golden_dir, cfg):
    data = np.load(os.path.join(golden_dir, cfg["name"] + ".npz"))
    pool = logits_pool(int(data["pool_seed"]), cfg["T"], cfg["V"], cfg["scale"])
    assert (zlib.crc32(pool.tobytes()) & 0xFFFFFFFF) == int(data["pool_crc"]), "input generator drifted"
    return data, pool


def test_ac_oracle_matches_reference_goldens(golden_dir, cases):
    for cfg in cases["ac"]:
        if cfg["V"] > 4096 and cfg["name"] != "ac_v50257_p26_full_t10":
            continue                                           # keep the CPU suite short
        data, pool = _load(golden_dir, cfg)
        kw = dict(temp=cfg["temp"], precision=cfg["precision"], topk=cfg["topk"])
        for s in range(min(cfg["streams"], 3)):
            rows = rows_for(pool, s)
            msg = data["msg_%d" % s].tolist()
            res = O.encode_stream(rows, msg, **kw)
            assert res.tokens == data["tokens_%d" % s].tolist(), (cfg["name"], s)
            tr = np.asarray([[t.new_bottom, t.new_top, t.nbits, t.lo, t.hi, t.k, t.selection] for t in res.trace])
            assert np.array_equal(tr, data["trace_%d" % s])
            bits, _ = O.decode_stream(rows, res.tokens, **kw)
            assert bits == data["decoded_%d" % s].tolist()
            assert bits[: len(msg)] == msg                     # round trip recovers the message


def test_codec_oracles_match_reference_goldens(golden_dir, cases):
    for cfg in cases["codecs"]:
        if cfg["V"] > 4096:
            continue
        data, pool = _load(golden_dir, cfg)
        for s in range(cfg["streams"]):
            rows = rows_for(pool, s)
            msg = data["msg_%d" % s].tolist()
            want_tok = data["tokens_%d" % s].tolist()
            want_bits = data["decoded_%d" % s].tolist()
            if cfg["kind"] == "huffman":
                tok, _ = K.huffman_encode(rows, msg, cfg["param"])
                assert tok == want_tok
                assert K.huffman_decode(rows, tok, cfg["param"]) == want_bits
            elif cfg["kind"] == "bins":
                tok, _ = K.bins_encode(rows, msg, cfg["param"], cfg["V"])
                assert tok == want_tok
                assert K.bins_decode(tok, cfg["param"], cfg["V"]) == want_bits
            else:
                payload = K.bits_to_bytes_msb(msg)
                top_k = cfg["param"] or None
                extra = K.rank_quality(cfg)                     # top_p / min_prob / cap_per_token_bits cases
                tok, hist, total = K.rank_encode(rows, payload, temperature=cfg["temperature"], top_k=top_k, **extra)
                assert tok == want_tok
                assert hist == data["history_%d" % s].tolist()
                back = K.rank_decode(rows, tok, hist, total, temperature=cfg["temperature"], top_k=top_k, **extra)
                assert K.bytes_to_bits_msb(back) == want_bits


def test_empty_and_tiny_messages():
    pool = logits_pool(5, 4, 2048, 3.0)
    res = O.encode_stream(rows_for(pool, 0), [], precision=16)
    assert res.tokens == [] and res.bits_consumed == 0
    res = O.encode_stream(rows_for(pool, 0), [1], precision=16)
    bits, _ = O.decode_stream(rows_for(pool, 0), res.tokens, precision=16)
    assert bits[:1] == [1]


def test_tie_break_is_lower_id_first():
    row = np.zeros(2048, dtype=np.float32)
    row[[7, 3, 900]] = 5.0
    s, order = O.sort_desc(row)
    assert order[:3].tolist() == [3, 7, 900]
