"""Framing around the coder: chunking, CRC32, Reed-Solomon, JSON packets, and the chunk -> stream batch.

Host-side mirror of the reference's framing layer, written so that every chunk of a message becomes one
stream of ONE batched coder call (the reference codes chunk after chunk, ``api.py:736-747``):

* ``crc32`` / ``append_crc32`` / ``verify_crc32``   -- ``framing/crc.py:12-45``, ``codec/packet.py:39-51``
* ``RSCodec`` / ``rs_encode`` / ``rs_decode``       -- the ``reedsolo`` package the reference imports
  (``codec/packet.py:15,54-65``, ``framing/ecc.py:23-51``; not installed in this image): systematic
  Reed-Solomon over GF(2^8), primitive polynomial 0x11d, generator 2, first consecutive root 0, messages
  cut into blocks of ``255 - nsym`` bytes.  Codewords are byte-identical to reedsolo's (systematic encoding
  with a fixed generator polynomial is unique); decoding corrects up to ``nsym // 2`` byte errors per block.
* ``build_packet`` / ``parse_packet``               -- ``codec/packet.py:68-160`` (byte-identical JSON)
* ``chunk_bytes`` / ``assemble_bytes`` / ``make_msg_id`` -- ``codec/chunker.py:8-40``
* ``stego_encode`` / ``stego_decode``               -- ``api.py:707-807`` with the chunks batched
"""

from __future__ import annotations

import base64
import json
import struct
import zlib
from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple
from uuid import uuid4

from .exceptions import ConfigurationError, MissingChunksError, PacketCRCError, PacketECCError

# ------------------------------------------------------------------------------------------ CRC32


def crc32(data: bytes) -> int:
    """zlib CRC32 (polynomial 0xEDB88320), unsigned (``framing/crc.py:12-20``)."""
    return zlib.crc32(data, 0) & 0xFFFFFFFF


def append_crc32(payload: bytes) -> bytes:
    """``payload`` + big-endian CRC32 (``codec/packet.py:39-41``)."""
    return payload + struct.pack(">I", crc32(payload))


def verify_crc32(blob: bytes) -> Tuple[bool, bytes]:
    """``(ok, payload_without_crc)`` (``framing/crc.py:31-45``)."""
    if len(blob) < 4:
        return False, blob
    payload, tail = blob[:-4], blob[-4:]
    return crc32(payload) == struct.unpack(">I", tail)[0], payload


# ------------------------------------------------------------------------------------------ GF(2^8) / Reed-Solomon
_PRIM = 0x11D
_EXP = [0] * 512
_LOG = [0] * 256
_x = 1
for _i in range(255):
    _EXP[_i] = _x
    _LOG[_x] = _i
    _x <<= 1
    if _x & 0x100:
        _x ^= _PRIM
for _i in range(255, 512):
    _EXP[_i] = _EXP[_i - 255]


def _mul(a: int, b: int) -> int:
    return 0 if a == 0 or b == 0 else _EXP[_LOG[a] + _LOG[b]]


def _div(a: int, b: int) -> int:
    if b == 0:
        raise ZeroDivisionError
    return 0 if a == 0 else _EXP[(_LOG[a] + 255 - _LOG[b]) % 255]


def _poly_eval(poly: Sequence[int], x: int) -> int:
    """Horner, coefficients highest degree first."""
    y = 0
    for c in poly:
        y = _mul(y, x) ^ c
    return y


def _poly_mul(p: Sequence[int], q: Sequence[int]) -> List[int]:
    r = [0] * (len(p) + len(q) - 1)
    for i, a in enumerate(p):
        if a:
            for j, b in enumerate(q):
                r[i + j] ^= _mul(a, b)
    return r


class ReedSolomonError(Exception):
    """Raised when a block holds more errors than ``nsym // 2``."""


class RSCodec:
    """``reedsolo.RSCodec(nsym)``-compatible codec (defaults: nsize 255, fcr 0, prim 0x11d, generator 2)."""

    def __init__(self, nsym: int = 10, nsize: int = 255):
        if nsym <= 0 or nsym >= nsize or nsize > 255:
            raise ValueError("nsym must be in (0, nsize) and nsize <= 255")
        self.nsym, self.nsize = int(nsym), int(nsize)
        g = [1]
        for i in range(self.nsym):                 # g(x) = prod (x - alpha^i), roots alpha^0 .. alpha^(nsym-1)
            g = _poly_mul(g, [1, _EXP[i]])
        self._gen = g

    # -------------------------------------------------------------- encode
    def _encode_block(self, msg: bytes) -> bytes:
        rem = [0] * self.nsym                      # remainder of msg(x) * x^nsym divided by g(x)
        gen = self._gen
        for byte in msg:
            coef = byte ^ rem[0]
            rem = rem[1:] + [0]
            if coef:
                for j in range(self.nsym):
                    rem[j] ^= _mul(gen[j + 1], coef)
        return bytes(msg) + bytes(rem)

    def encode(self, data) -> bytearray:
        data = bytes(data)
        k = self.nsize - self.nsym
        out = bytearray()
        for i in range(0, len(data), k):
            out += self._encode_block(data[i:i + k])
        return out

    # -------------------------------------------------------------- decode
    def _syndromes(self, block: Sequence[int]) -> List[int]:
        return [_poly_eval(block, _EXP[i]) for i in range(self.nsym)]

    def _correct_block(self, block: bytearray) -> Tuple[bytearray, List[int]]:
        synd = self._syndromes(block)
        if not any(synd):
            return block, []
        # Berlekamp-Massey: error locator sigma(x), lowest degree first
        sigma, prev = [1], [1]
        L, m, b = 0, 1, 1
        for n in range(self.nsym):
            d = synd[n]
            for i in range(1, L + 1):
                if i < len(sigma):
                    d ^= _mul(sigma[i], synd[n - i])
            if d == 0:
                m += 1
                continue
            coef = _div(d, b)
            shifted = [0] * m + [_mul(coef, c) for c in prev]
            new = [(sigma[i] if i < len(sigma) else 0) ^ (shifted[i] if i < len(shifted) else 0)
                   for i in range(max(len(sigma), len(shifted)))]
            if 2 * L <= n:
                prev, L, b, m = sigma, n + 1 - L, d, 1
            else:
                m += 1
            sigma = new
        while len(sigma) > 1 and sigma[-1] == 0:
            sigma.pop()
        nerr = len(sigma) - 1
        if nerr != L or 2 * nerr > self.nsym:
            raise ReedSolomonError("Too many errors to correct")
        # Chien search: position p (from the end, 0 = last byte) is in error iff sigma(alpha^-p) == 0
        n = len(block)
        positions = []
        for p in range(n):
            xinv = _EXP[(255 - p) % 255]
            v = 0
            for c in reversed(sigma):
                v = _mul(v, xinv) ^ c
            if v == 0:
                positions.append(p)
        if len(positions) != nerr:
            raise ReedSolomonError("Too many (or few) errors found by Chien Search for the errata locator polynomial!")
        # Forney: omega(x) = S(x) sigma(x) mod x^nsym (lowest first), e = X omega(X^-1) / sigma'(X^-1)   (fcr = 0)
        omega = [0] * self.nsym
        for i, s in enumerate(synd):
            for j, c in enumerate(sigma):
                if i + j < self.nsym:
                    omega[i + j] ^= _mul(s, c)
        errata = []
        for p in positions:
            X = _EXP[p % 255]
            xinv = _EXP[(255 - p) % 255]
            num = 0
            for c in reversed(omega):
                num = _mul(num, xinv) ^ c
            den = 0
            for i in range(1, len(sigma), 2):      # formal derivative: odd-degree terms
                den ^= _mul(sigma[i], _EXP[(_LOG[xinv] * (i - 1)) % 255] if xinv else 0)
            if den == 0:
                raise ReedSolomonError("Could not find error magnitude")
            mag = _mul(X, _div(num, den))
            block[n - 1 - p] ^= mag
            errata.append(n - 1 - p)
        if any(self._syndromes(block)):
            raise ReedSolomonError("Could not correct message")
        return block, errata

    def decode(self, data) -> Tuple[bytearray, bytearray, List[int]]:
        """-> (message, message + ecc, errata positions), like reedsolo >= 1.5."""
        data = bytearray(data)
        msg, full, errata_all = bytearray(), bytearray(), []
        for i in range(0, len(data), self.nsize):
            block = bytearray(data[i:i + self.nsize])
            if len(block) <= self.nsym:
                raise ReedSolomonError("block shorter than the parity")
            fixed, errata = self._correct_block(block)
            msg += fixed[:-self.nsym]
            full += fixed
            errata_all += [i + e for e in errata]
        return msg, full, errata_all


def rs_encode(data: bytes, nsym: int = 10) -> bytes:
    """``framing/ecc.py:23-27``."""
    return bytes(RSCodec(nsym).encode(bytearray(data)))


def rs_decode(codeword: bytes, nsym: int = 10) -> Tuple[bool, bytes]:
    """``(ok, data)``; ``(False, b"")`` when a block cannot be corrected (``framing/ecc.py:30-51``)."""
    try:
        msg, _, _ = RSCodec(nsym).decode(bytearray(codeword))
    except ReedSolomonError:
        return False, b""
    return True, bytes(msg)


# ------------------------------------------------------------------------------------------ chunks and packets
def make_msg_id() -> str:
    return str(uuid4())                                                   # codec/chunker.py:8-11


def chunk_bytes(data: bytes, *, chunk_size: int = 256) -> List[bytes]:
    if chunk_size <= 0:
        raise ValueError("chunk_size must be positive")                   # codec/chunker.py:28-29
    if not data:
        return [b""]
    return [data[i:i + chunk_size] for i in range(0, len(data), chunk_size)]


def assemble_bytes(chunks: Iterable[bytes]) -> bytes:
    return b"".join(chunks)


@dataclass(frozen=True)
class Packet:
    msg_id: str
    seq: int
    total: int
    cfg: Dict[str, Any]
    payload: bytes


def _cfg_norm(cfg: Mapping[str, Any]) -> Dict[str, Any]:
    return {"chunk_bytes": cfg.get("chunk_bytes"), "crc": bool(cfg.get("crc", False)),
            "ecc": cfg.get("ecc", "none"), "nsym": int(cfg.get("nsym", 0))}


def build_packet(payload: bytes, *, msg_id: str, seq: int, total: int, cfg: Mapping[str, Any]) -> bytes:
    """Byte-identical to ``codec/packet.py:68-106``: CRC32 append, RS encode, base64, sorted-key compact JSON."""
    if seq < 0 or total <= 0 or seq >= total:
        raise ValueError("invalid sequence/total combination")
    c = _cfg_norm(cfg)
    framed = payload
    if c["crc"]:
        framed = append_crc32(framed)
    if c["ecc"] == "rs":
        if c["nsym"] <= 0:
            raise ValueError("nsym must be positive when ecc='rs'")
        framed = rs_encode(framed, c["nsym"])
    elif c["ecc"] not in {"none", None}:
        raise ConfigurationError("unsupported ecc mode: %s" % c["ecc"])
    obj = {"version": 1, "msg_id": msg_id, "seq": seq, "total": total, "cfg": c,
           "payload": base64.b64encode(framed).decode("ascii")}
    return json.dumps(obj, separators=(",", ":"), sort_keys=True).encode("utf-8")


def parse_packet(packet: bytes, *, expected_cfg: Optional[Mapping[str, Any]] = None) -> Packet:
    """``codec/packet.py:109-160``: JSON -> base64 -> RS decode -> CRC verify."""
    try:
        obj = json.loads(packet.decode("utf-8"))
    except (ValueError, UnicodeDecodeError) as exc:
        raise PacketECCError("invalid packet encoding") from exc
    required = {"msg_id", "seq", "total", "cfg", "payload"}
    if not isinstance(obj, dict) or not required.issubset(obj):
        missing = ", ".join(sorted(required - set(obj))) if isinstance(obj, dict) else "all"
        raise PacketECCError("missing packet keys: %s" % missing)
    c = _cfg_norm(obj["cfg"])
    if expected_cfg:
        for key, value in expected_cfg.items():
            if key in c and value is not None and c[key] != value:
                raise ConfigurationError("packet cfg mismatch for %s: expected %s, got %s" % (key, value, c[key]))
    try:
        framed = base64.b64decode(obj["payload"], validate=True)
    except (ValueError, TypeError) as exc:
        raise PacketECCError("payload is not valid base64") from exc
    if c["ecc"] == "rs":
        ok, framed = rs_decode(framed, c["nsym"])
        if not ok:
            raise PacketECCError("Reed-Solomon decoding failed")
    elif c["ecc"] not in {"none", None}:
        raise ConfigurationError("unsupported ecc mode: %s" % c["ecc"])
    if c["crc"]:
        if len(framed) < 4:
            raise PacketCRCError("payload too small to contain CRC32")
        ok, framed = verify_crc32(framed)
        if not ok:
            raise PacketCRCError("CRC32 mismatch detected")
    return Packet(msg_id=obj["msg_id"], seq=int(obj["seq"]), total=int(obj["total"]), cfg=c, payload=framed)


# ------------------------------------------------------------------------------------------ chunk batch <-> provider
_DEFAULT_QUALITY = {"temp": 1.0, "precision": 16, "topk": 50000, "finish_sent": True}        # api.py:81-86
_QUALITY_KEY_ALIASES = {                                                                      # api.py:130-141
    "temperature": "temp", "top-k": "top_k", "topk": "top_k", "top_p": "top_p", "top-p": "top_p",
    "cap-per-token-bits": "cap_per_token_bits", "cap_bits_per_token": "cap_per_token_bits",
    "cap-bits-per-token": "cap_per_token_bits", "max-context": "max_context", "maxContext": "max_context"}


def normalise_quality_dict(quality: Optional[Mapping[str, object]]) -> Dict[str, Any]:
    return {_QUALITY_KEY_ALIASES.get(k, k): v for k, v in (quality or {}).items()}            # api.py:180-189


def bytes_to_bits(data: bytes) -> List[int]:
    return [(byte >> i) & 1 for byte in data for i in range(8)]                               # api.py:153-157


def bits_to_bytes(bits: Iterable[int]) -> bytes:
    bl = list(bits)
    if len(bl) % 8:
        raise ConfigurationError("decoded bit stream is not byte aligned")                    # api.py:162-163
    out = bytearray()
    for i in range(0, len(bl), 8):
        v = 0
        for off, bit in enumerate(bl[i:i + 8]):
            v |= (bit & 1) << off
        out.append(v)
    return bytes(out)


@dataclass
class EncodeMetadata:
    msg_id: str
    total: int
    cfg: Dict[str, object]


class EncodeResult(list):
    """List of token spans with the framing metadata attached (``api.py:66-71``)."""

    def __init__(self, spans: Iterable[List[int]], metadata: EncodeMetadata) -> None:
        super().__init__(spans)
        self.metadata = metadata


def _normalise_ecc(ecc: Optional[str]) -> str:
    if not ecc:
        return "none"
    e = ecc.lower()
    if e not in {"none", "rs"}:
        raise ConfigurationError("unsupported ecc mode: %s" % ecc)                            # api.py:146-149
    return e


def stego_encode(message: bytes, *, chunk_bytes: int = 256, use_crc: bool = True, ecc: Optional[str] = "rs",
                 nsym: int = 10, quality: Optional[Mapping[str, object]] = None, seed_text: str = "", lm,
                 msg_id: Optional[str] = None) -> EncodeResult:
    """``api.stego_encode`` (``api.py:707-749``) with all chunks of the message coded as ONE stream batch when the
    provider offers ``encode_arithmetic_batch`` (the device provider does); otherwise chunk after chunk as in the
    reference.  ``msg_id`` may be injected (the reference draws a random UUID)."""
    ecc_mode = _normalise_ecc(ecc)
    cfg = {"chunk_bytes": int(chunk_bytes), "crc": bool(use_crc), "ecc": ecc_mode,
           "nsym": int(nsym if ecc_mode == "rs" else 0)}
    quality_args = {**_DEFAULT_QUALITY, **normalise_quality_dict(quality)}
    chunks = _chunk(message, cfg["chunk_bytes"])           # (the keyword argument shadows the function's name here)
    mid = msg_id or make_msg_id()
    total = len(chunks)
    packets = [build_packet(ch, msg_id=mid, seq=seq, total=total, cfg=cfg) for seq, ch in enumerate(chunks)]
    bit_lists = [bytes_to_bits(p) for p in packets]
    context = lm.encode_seed(seed_text)
    if hasattr(lm, "encode_arithmetic_batch"):
        spans = lm.encode_arithmetic_batch(bit_lists, context, quality=quality_args)
    else:
        spans = [lm.encode_arithmetic(bits, context, quality=quality_args) for bits in bit_lists]
    return EncodeResult([list(s) for s in spans], EncodeMetadata(msg_id=mid, total=total, cfg=cfg))


def _chunk(message: bytes, size: int) -> List[bytes]:
    return chunk_bytes(message, chunk_size=size)


def stego_decode(spans: Iterable[Sequence[int]], *, use_crc: bool = True, ecc: Optional[str] = "rs", nsym: int = 10,
                 quality: Optional[Mapping[str, object]] = None, seed_text: str = "", lm) -> bytes:
    """``api.stego_decode`` (``api.py:752-807``), spans decoded as one stream batch when the provider can."""
    ecc_mode = _normalise_ecc(ecc)
    expected_cfg = {"crc": bool(use_crc), "ecc": ecc_mode, "nsym": int(nsym if ecc_mode == "rs" else 0)}
    quality_args = {**_DEFAULT_QUALITY, **normalise_quality_dict(quality)}
    span_lists = [list(s) for s in spans]
    context = lm.encode_seed(seed_text)
    if span_lists and hasattr(lm, "decode_arithmetic_batch"):
        bit_lists = lm.decode_arithmetic_batch(span_lists, context, quality=quality_args)
    else:
        bit_lists = [lm.decode_arithmetic(s, context, quality=quality_args) for s in span_lists]
    payload_by_seq: Dict[int, bytes] = {}
    msg_id: Optional[str] = None
    total: Optional[int] = None
    for bits in bit_lists:
        packet = parse_packet(bits_to_bytes(bits), expected_cfg=expected_cfg)
        if msg_id is None:
            msg_id, total = packet.msg_id, packet.total
        else:
            if packet.msg_id != msg_id:
                raise ConfigurationError("decoded packet msg_id mismatch")
            if packet.total != total:
                raise ConfigurationError("decoded packet total mismatch")
        if packet.seq in payload_by_seq:
            raise ConfigurationError("duplicate packet sequence %d" % packet.seq)
        payload_by_seq[packet.seq] = packet.payload
    if total is None:
        return b""
    present = sorted(payload_by_seq)
    missing = sorted(set(range(total)) - set(payload_by_seq))
    assembled = assemble_bytes([payload_by_seq[i] for i in present])
    if missing:
        raise MissingChunksError(missing_indices=missing, partial_payload=assembled)
    return assembled


__all__ = ["crc32", "append_crc32", "verify_crc32", "RSCodec", "ReedSolomonError", "rs_encode", "rs_decode",
           "make_msg_id", "chunk_bytes", "assemble_bytes", "Packet", "build_packet", "parse_packet",
           "bytes_to_bits", "bits_to_bytes", "normalise_quality_dict", "EncodeMetadata", "EncodeResult",
           "stego_encode", "stego_decode"]
