"""Framing layer (host, no GPU): CRC32, Reed-Solomon, packets, chunk batches.

Known answers: the reference's own tests (tests/framing/test_crc.py:4-6, test_ecc.py:6-48, test_packet.py,
test_api_end_to_end.py:9-46, test_fault_injection.py:39-110), reedsolo's published example, and packets produced by the
reference's build_packet (tests/golden/framing_packets.json, oracle/make_framing_golden.py)."""
import base64
import json
import os
import random

import pytest

from neuralsteganography_b200 import framing as F
from neuralsteganography_b200.exceptions import ConfigurationError, MissingChunksError, PacketCRCError, PacketECCError
from neuralsteganography_b200.lm import MockLM


def test_crc32_known_value_and_detection():
    assert F.crc32(b"hello") == 0x3610A686                      # tests/framing/test_crc.py:4-6
    blob = F.append_crc32(b"payload")
    assert F.verify_crc32(blob) == (True, b"payload")
    bad = bytearray(blob); bad[0] ^= 0xFF
    ok, rec = F.verify_crc32(bytes(bad))
    assert not ok and rec != b"payload"
    assert F.verify_crc32(b"abc") == (False, b"abc")


def test_rs_matches_reedsolo_published_codeword():
    # reedsolo README: RSCodec(10).encode(b'hello world')
    assert bytes(F.RSCodec(10).encode(b"hello world")) == b"hello world\xed%T\xc4\xfd\xfd\x89\xf3\xa8\xaa"


def test_rs_roundtrip_correction_and_failure():
    ok, dec = F.rs_decode(F.rs_encode(b"neural stego", nsym=8), nsym=8)      # test_ecc.py:6-14
    assert ok and dec == b"neural stego"
    enc = bytearray(F.rs_encode(b"0123456789abcdef", nsym=8))                # test_ecc.py:17-32
    enc[0] ^= 0x01; enc[3] ^= 0x01; enc[5] ^= 0x02
    assert F.rs_decode(bytes(enc), nsym=8) == (True, b"0123456789abcdef")
    enc = bytearray(F.rs_encode(b"another block", nsym=4))                   # test_ecc.py:35-48
    enc[0] ^= 0x01; enc[1] ^= 0x02; enc[2] ^= 0x04
    assert F.rs_decode(bytes(enc), nsym=4) == (False, b"")


def test_rs_corrects_up_to_half_nsym_anywhere_and_multi_block():
    rng = random.Random(5)
    for nsym in (2, 4, 10, 32):
        data = bytes(rng.randrange(256) for _ in range(700))                 # three blocks at nsym <= 10
        enc = bytearray(F.rs_encode(data, nsym))
        assert len(enc) == len(data) + nsym * -(-len(data) // (255 - nsym))
        for blk in range(0, len(enc), 255):
            hi = min(blk + 255, len(enc))
            for pos in rng.sample(range(blk, hi), nsym // 2):
                enc[pos] ^= rng.randrange(1, 256)
        assert F.rs_decode(bytes(enc), nsym) == (True, data)
    msg, full, errata = F.RSCodec(4).decode(F.RSCodec(4).encode(b"abc"))
    assert bytes(msg) == b"abc" and errata == []


def test_packets_are_byte_identical_to_the_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "framing_packets.json")))
    assert len(gold["cases"]) == 24
    for c in gold["cases"]:
        payload = base64.b64decode(c["payload_b64"])
        pkt = F.build_packet(payload, msg_id=gold["msg_id"], seq=c["seq"], total=c["total"], cfg=c["cfg"])
        assert pkt.decode("utf-8") == c["packet"]
        back = F.parse_packet(pkt, expected_cfg={k: c["cfg"][k] for k in ("crc", "ecc", "nsym")})
        assert back.payload == payload and back.seq == c["seq"] and back.msg_id == gold["msg_id"]


def test_packet_errors():
    cfg = {"chunk_bytes": 64, "crc": True, "ecc": "rs", "nsym": 10}
    pkt = F.build_packet(b"abc", msg_id="m", seq=0, total=1, cfg=cfg)
    with pytest.raises(ValueError):
        F.build_packet(b"", msg_id="m", seq=1, total=1, cfg=cfg)             # codec/packet.py:78-79
    with pytest.raises(ConfigurationError):
        F.parse_packet(pkt, expected_cfg={"nsym": 4})
    with pytest.raises(PacketECCError):
        F.parse_packet(b"not json")
    obj = json.loads(pkt); obj["cfg"]["ecc"] = "none"                         # RS parity now looks like payload: CRC must catch it
    with pytest.raises(PacketCRCError):
        F.parse_packet(json.dumps(obj).encode())


@pytest.mark.parametrize("use_crc,ecc", [(True, "rs"), (True, "none"), (False, "none"), (False, "rs")])
def test_stego_roundtrip_over_the_mock_provider(use_crc, ecc):
    # tests/framing/test_api_end_to_end.py:9-46 (config 1: --model mock, CRC/ECC on and off)
    rng = random.Random(11)
    message = bytes(rng.randrange(256) for _ in range(4096))
    lm = MockLM()
    res = F.stego_encode(message, chunk_bytes=256, use_crc=use_crc, ecc=ecc, nsym=10, seed_text="seed", lm=lm)
    assert res.metadata.total == 16 and len(res) == 16
    out = F.stego_decode(res, use_crc=use_crc, ecc=ecc, nsym=10, seed_text="seed", lm=lm)
    assert out == message
    assert F.stego_decode(F.stego_encode(b"", lm=lm), lm=lm) == b""


def test_fault_injection_rs_repairs_and_missing_chunk_is_reported():
    # tests/framing/test_fault_injection.py:39-110
    lm = MockLM()
    message = bytes(range(200)) * 3
    res = F.stego_encode(message, chunk_bytes=128, use_crc=True, ecc="rs", nsym=10, lm=lm, msg_id="fixed")
    spans = [list(s) for s in res]
    pkt = bytes(spans[1])
    obj = json.loads(pkt)
    b64 = list(obj["payload"])
    b64[5] = "A" if b64[5] != "A" else "B"                      # flip one base64 symbol
    obj["payload"] = "".join(b64)
    spans[1] = list(json.dumps(obj, separators=(",", ":"), sort_keys=True).encode())
    assert F.stego_decode(spans, use_crc=True, ecc="rs", nsym=10, lm=lm) == message
    with pytest.raises(MissingChunksError) as ei:
        F.stego_decode(spans[:2] + spans[3:], use_crc=True, ecc="rs", nsym=10, lm=lm)
    assert ei.value.missing_indices == [2] and ei.value.partial_payload == message[:256] + message[384:]


def test_batched_provider_receives_all_chunks_in_one_call():
    calls = []

    class Batch(MockLM):
        def encode_arithmetic_batch(self, bit_lists, context, *, quality):
            calls.append(len(bit_lists))
            return [self.encode_arithmetic(b, context, quality=quality) for b in bit_lists]

        def decode_arithmetic_batch(self, token_lists, context, *, quality):
            calls.append(-len(token_lists))
            return [self.decode_arithmetic(t, context, quality=quality) for t in token_lists]

    lm = Batch()
    res = F.stego_encode(b"x" * 1000, chunk_bytes=100, lm=lm)
    assert F.stego_decode(res, lm=lm) == b"x" * 1000
    assert calls == [10, -10]
