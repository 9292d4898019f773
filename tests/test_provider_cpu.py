"""Host logic of the provider boundary (no GPU): mock provider, quality normalisation, packet trimming,
error behaviour, and -- where the reference is mounted -- its own stego_encode/stego_decode on top of it."""
import json
import os
import sys

import pytest

from neuralsteganography_b200 import lm as L
from neuralsteganography_b200.exceptions import ConfigurationError


def test_mock_provider_is_byte_identity():
    m = L.load_lm("mock")
    bits = L.bytes_to_bits_lsb(b"hello \xff\x00")
    toks = m.encode_arithmetic(bits, [1, 2], quality={})
    assert toks == list(b"hello \xff\x00")                      # lm/mock.py:45-52
    assert m.decode_arithmetic(toks, [1, 2], quality={}) == bits
    assert m.encode_arithmetic([], [1], quality={}) == []
    assert m.encode_seed("ab") == [97, 98]


def test_bit_glue_is_lsb_first_and_checks_alignment():
    assert L.bytes_to_bits_lsb(b"\x01") == [1, 0, 0, 0, 0, 0, 0, 0]     # api.py:153-157
    assert L.bits_to_bytes_lsb([0, 1, 0, 0, 0, 0, 0, 0]) == b"\x02"
    with pytest.raises(ConfigurationError):
        L.bits_to_bytes_lsb([1, 0, 1])                           # lm/arithmetic.py:18-19


def test_quality_aliases_and_validation():
    q = L.normalise_quality({"temperature": 0.9, "top-k": "300", "precision": 26, "finish_sent": "false"})
    assert q == {"temp": 0.9, "precision": 26, "topk": 300, "finish_sent": False}
    assert L.normalise_quality(None) == {"temp": 1.0, "precision": 16, "topk": 50000, "finish_sent": True}   # api.py:81-86
    assert isinstance(L.normalise_quality({"topk": 50000.0})["topk"], int)     # the reference's float top_k bug
    for bad in ({"temp": 0}, {"precision": 99}, {"topk": 0}):
        with pytest.raises(ConfigurationError):
            L.normalise_quality(bad)


def test_trim_to_packet_cuts_trailing_coder_bits():
    pkt = json.dumps({"cfg": {"crc": True, "ecc": "none"}, "payload": "aGVsbG8=", "seq": 0},
                     separators=(",", ":"), sort_keys=True).encode()
    bits = L.bytes_to_bits_lsb(pkt) + [1, 0, 1, 1, 0, 0, 1, 0, 1, 1, 1]
    assert L.bits_to_bytes_lsb(L._trim_to_packet(bits)) == pkt
    assert len(L._trim_to_packet([1] * 13)) == 8                 # not a packet: whole bytes only


def test_unknown_provider_and_missing_gpu_fail_loudly():
    with pytest.raises(ConfigurationError):
        L.load_lm("no-such-model")                               # lm/__init__.py:26
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(ConfigurationError):
            L.B200ArithmeticLM(_Dummy(), L.IdTokenizer(10))


class _Dummy:
    def eval(self):
        return self


def test_id_tokenizer_roundtrip():
    t = L.IdTokenizer(50257)
    assert t.encode(t.decode([5, 17, 50000])) == [5, 17, 50000]
    assert t.encode("<|endoftext|>") == [50256]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/neuralstego"), reason="reference not mounted")
def test_reference_pipeline_accepts_the_mock_provider():
    """config 1: the reference's own stego_encode/stego_decode (api.py:707-807) driven by this provider."""
    sys.path.insert(0, "/root/reference/src")
    sys.path.insert(0, "/root/reference")
    try:
        from neuralstego.api import stego_decode, stego_encode
    except Exception as exc:  # pragma: no cover
        pytest.skip("reference api not importable: %s" % exc)
    lm = L.MockLM()
    payload = bytes(range(256)) * 3
    for use_crc in (True, False):
        spans = stego_encode(payload, chunk_bytes=200, use_crc=use_crc, ecc="none", quality={"temp": 0.9}, seed_text="x", lm=lm)
        assert len(spans) == 4
        assert stego_decode(spans, use_crc=use_crc, ecc="none", quality={"temp": 0.9}, seed_text="x", lm=lm) == payload
