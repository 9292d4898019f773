"""Provider boundary: the reference's ``LMProvider`` protocol on top of the device coder.

Mirrors ``src/neuralstego/api.py:42-56`` (protocol), ``lm/__init__.py:16-26`` (``load_lm``),
``lm/arithmetic.py:115-235`` (``ArithmeticLM``) and ``lm/mock.py:36-59`` (``MockLM``).
``stego_encode`` / ``stego_decode`` (api.py:707-807) can be handed a :class:`B200ArithmeticLM`
unchanged: same method names, argument meaning and error type.

What is different on purpose (each is a defect of the reference):
  * the codec is the finite-precision arithmetic coder of code_base/arithmetic.py -- what the
    API's default quality dict ``{temp, precision, topk, finish_sent}`` (api.py:81-86) is written
    for -- not the fixed-width rank coder; no side-channel ``history`` is needed to decode;
  * ``topk`` stays an integer (the reference turns it into a float and crashes for V > topk);
  * the bit list returned by ``decode_arithmetic`` is cut after the framed packet's closing brace
    so that ``api._bits_to_bytes`` / ``parse_packet`` receive exactly the packet.
"""

from __future__ import annotations

from typing import Dict, Iterable, List, Mapping, Optional, Sequence

import torch

from .exceptions import ConfigurationError

_MODEL_ALIASES = {"gpt2-fa": "HooshvareLab/gpt2-fa"}          # lm/__init__.py:11-13
_DEFAULT_QUALITY = {"temp": 1.0, "precision": 16, "topk": 50000, "finish_sent": True}   # api.py:81-86


# ---------------------------------------------------------------------------- bit glue (api.py:153-172)
def bits_to_bytes_lsb(bits: Iterable[int]) -> bytes:
    data = [int(b) & 1 for b in bits]
    if len(data) % 8 != 0:
        raise ConfigurationError("bit stream length must be a multiple of 8")     # lm/arithmetic.py:18-19
    out = bytearray()
    for i in range(0, len(data), 8):
        v = 0
        for off, bit in enumerate(data[i:i + 8]):
            v |= bit << off
        out.append(v)
    return bytes(out)


def bytes_to_bits_lsb(payload: bytes) -> List[int]:
    return [(byte >> off) & 1 for byte in payload for off in range(8)]


def normalise_quality(quality: Optional[Mapping[str, object]]) -> Dict[str, object]:
    """Overlay user keys on the API defaults, accepting the reference's aliases (api.py:130-141)."""
    q: Dict[str, object] = dict(_DEFAULT_QUALITY)
    for key, value in (quality or {}).items():
        if value is None:
            continue
        k = str(key).replace("-", "_").lower()
        if k in ("temp", "temperature"):
            q["temp"] = float(value)
        elif k in ("topk", "top_k"):
            q["topk"] = int(float(value))
        elif k == "precision":
            q["precision"] = int(value)
        elif k == "finish_sent":
            q["finish_sent"] = bool(value) if not isinstance(value, str) else value.lower() in ("1", "true", "yes")
        elif k in ("topp", "top_p"):                          # lm/arithmetic.py:88-93, api.py:130-141
            q["top_p"] = float(value)
        elif k in ("minprob", "min_prob"):
            q["min_prob"] = float(value)
        elif k in ("cap_per_token_bits", "cap_bits_per_token"):
            q["cap_per_token_bits"] = int(float(value))
        elif k in ("max_context", "maxcontext"):
            q["max_context"] = int(value)
        else:
            q[k] = value
    if not (q["temp"] > 0):
        raise ConfigurationError("temperature must be positive")                  # lm/arithmetic.py:70-71
    if not (2 <= int(q["precision"]) <= 48):
        raise ConfigurationError("precision must be in [2, 48]")
    if int(q["topk"]) < 1:
        raise ConfigurationError("topk must be >= 1")
    return q


def _trim_to_packet(bits: List[int]) -> List[int]:
    """Cut the recovered bits after the JSON packet (codec/packet.py:97-106) and to whole bytes."""
    usable = len(bits) - len(bits) % 8
    data = bits_to_bytes_lsb(bits[:usable])
    if data[:1] == b"{":
        end = data.find(b"}")
        while end != -1:                      # the cfg object nests one level: take the brace that balances
            if data[:end + 1].count(b"{") == data[:end + 1].count(b"}"):
                return bits[: (end + 1) * 8]
            end = data.find(b"}", end + 1)
    return bits[:usable]


# ---------------------------------------------------------------------------- tokenizers / mock
class IdTokenizer:
    """Offline stand-in tokenizer: text is a space-separated list of token ids."""

    def __init__(self, vocab_size: int):
        self.vocab_size = int(vocab_size)

    def encode(self, text: str, add_special_tokens: bool = False) -> List[int]:
        if text == "<|endoftext|>":
            return [self.vocab_size - 1]
        out = []
        for piece in text.split():
            out.append(int(piece) % self.vocab_size if piece.lstrip("-").isdigit() else sum(piece.encode("utf-8")) % self.vocab_size)
        return out

    def decode(self, ids: Iterable[int], skip_special_tokens: bool = True) -> str:
        return " ".join(str(int(i)) for i in ids)


class MockTokenizer:
    """lm/mock.py:9-14."""

    def encode(self, text: str) -> List[int]:
        return list(text.encode("utf-8"))

    def decode(self, tokens: Iterable[int]) -> str:
        return bytes(int(t) % 256 for t in tokens).decode("utf-8", errors="ignore")


class MockLM:
    """Byte-identity provider of ``--model mock`` (lm/mock.py:36-59): tokens are the packet bytes."""

    def __init__(self) -> None:
        self.tokenizer = MockTokenizer()

    def encode_seed(self, text: str) -> List[int]:
        return self.tokenizer.encode(text)

    def encode_arithmetic(self, bits: List[int], context: List[int], *, quality: Dict[str, float]) -> List[int]:
        if not bits:
            return []
        return [int(b) for b in bits_to_bytes_lsb(bits)]

    def decode_arithmetic(self, tokens: List[int], context: List[int], *, quality: Dict[str, float]) -> List[int]:
        return bytes_to_bits_lsb(bytes(int(t) % 256 for t in tokens))


# ---------------------------------------------------------------------------- the device provider
class B200ArithmeticLM:
    """``LMProvider`` whose coder runs on the GPU (one call = one stream; ``*_batch`` = many streams)."""

    codec = "ac"

    def __init__(self, model, tokenizer, *, device: Optional[str] = None, max_len: int = 1024, use_graph: bool = True,
                 batch_size: Optional[int] = None, trunk_tf32: bool = False):
        """``batch_size`` pins the number of streams every trunk call runs with (shorter batches are padded with idle
        streams).  Encoder and decoder must see bit-identical logits, and the trunk's cuBLAS GEMMs pick their
        algorithm -- hence the fp32 summation order -- by batch shape: a cover encoded as one of N streams has to be
        decoded at the same N.  Within one provider instance that is automatic (a decode is padded to the batch of
        the last encode); a process that only decodes passes the encoder's ``batch_size`` here.

        ``trunk_tf32`` runs the trunk's GEMMs (the ``lm_head`` projection above all) on the tensor cores in TF32: the
        logits then differ from an fp32 trunk's in their low bits, so encoder and decoder must both set it."""
        if not torch.cuda.is_available():
            raise ConfigurationError("%s needs a CUDA device (there is no CPU fallback)" % type(self).__name__)
        self.model = model.eval()
        self.tokenizer = tokenizer
        self.device = torch.device(device or "cuda")
        self.max_len = int(max_len)
        self.use_graph = bool(use_graph)
        self.batch_size = int(batch_size) if batch_size else None
        self.trunk_tf32 = bool(trunk_tf32)
        self._last_encode_batch: Optional[int] = None
        self._gens: Dict[tuple, object] = {}
        self._sent_end: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ protocol
    def encode_seed(self, text: str) -> List[int]:
        """``<|endoftext|>`` + seed ids (lm/arithmetic.py:143-160)."""
        tok = self.tokenizer
        if not hasattr(tok, "encode"):
            return list(text.encode("utf-8"))
        try:
            bos = list(tok.encode("<|endoftext|>", add_special_tokens=False))
        except TypeError:
            bos = list(tok.encode("<|endoftext|>"))
        except Exception:
            bos = []
        try:
            ids = list(tok.encode(text, add_special_tokens=False))
        except TypeError:
            ids = list(tok.encode(text))
        return [int(t) for t in bos + ids]

    def encode_arithmetic(self, bits: List[int], context: List[int], *, quality: Mapping[str, object]) -> List[int]:
        return self.encode_arithmetic_batch([bits], context, quality=quality)[0]

    def decode_arithmetic(self, tokens: List[int], context: List[int], *, quality: Mapping[str, object]) -> List[int]:
        return self.decode_arithmetic_batch([tokens], context, quality=quality)[0]

    def drain_states(self) -> list:          # api.py:849-854: the arithmetic coder needs no side information
        return []

    def load_states(self, states) -> None:   # api.py:996-1000
        return None

    # ------------------------------------------------------------------ batched streams
    def _reject_rank_policies(self, q: Mapping[str, object]) -> None:
        """``top_p`` / ``min_prob`` / ``cap_per_token_bits`` are policies of the rank codec (codec/quality.py:57-141);
        the arithmetic coder of code_base/arithmetic.py has no such knobs -- say so instead of ignoring them."""
        bad = [k for k in ("top_p", "min_prob", "cap_per_token_bits") if q.get(k) is not None]
        if bad:
            raise ConfigurationError("quality key(s) %s belong to the rank codec; use B200RankLM / load_lm(..., codec='rank') "
                                     "or drop them" % ", ".join(bad))

    def _batch_for(self, n: int, decode: bool) -> int:
        if self.batch_size:
            if n > self.batch_size:
                raise ConfigurationError("%d streams exceed the provider's pinned batch_size=%d" % (n, self.batch_size))
            return self.batch_size
        if decode and self._last_encode_batch and n <= self._last_encode_batch:
            return self._last_encode_batch           # same trunk shapes as the encode that produced the cover
        return n

    def _generator(self, batch: int, q: Mapping[str, object]):
        from .generation import StegoGenerator
        key = (batch, float(q["temp"]), int(q["precision"]), int(q["topk"]), bool(q["finish_sent"]))
        gen = self._gens.get(key)
        if gen is None:
            gen = StegoGenerator(self.model, batch, max_len=self.max_len, precision=int(q["precision"]),
                                 temp=float(q["temp"]), topk=int(q["topk"]), finish_sent=bool(q["finish_sent"]),
                                 sent_end=self._sentence_end_table() if q["finish_sent"] else None,
                                 device=self.device, use_graph=self.use_graph, trunk_tf32=self.trunk_tf32)
            self._gens = {key: gen}            # keep one (KV buffers are large)
        return gen

    def _raise_on_status(self, gen) -> None:
        bits = 0
        for v in gen.coder.status.unique().tolist():
            bits |= int(v)
        if bits & 8:
            raise ConfigurationError("cover did not fit %d tokens; raise max_len (>= 1023 slides the window) or shorten "
                                     "the chunk" % gen.coder.token_cap)
        if bits & 2:
            # a selection bucket overflowed (thousands of exactly equal logits): the order inside it was cut
            # arbitrarily, encoder and decoder are no longer guaranteed to agree
            raise ConfigurationError("degenerate logits (massive exact ties) overflowed the coder's selection bucket; "
                                     "the cover would not be decodable -- use a trunk with more numerical resolution")

    def _sentence_end_table(self) -> Optional[torch.Tensor]:
        """[V] 1 where the token text contains . ! ? (code_base/utils.py:55-57)."""
        if self._sent_end is None:
            V = self.model.config.vocab_size
            flags = torch.zeros(V, dtype=torch.uint8)
            for i in range(V):
                try:
                    s = self.tokenizer.decode([i])
                except Exception:
                    s = ""
                if "." in s or "!" in s or "?" in s:
                    flags[i] = 1
            self._sent_end = flags.to(self.device)
        return self._sent_end

    def _check_context(self, context: Sequence[int]) -> torch.Tensor:
        ids = [int(t) for t in context]
        if not ids:
            raise ConfigurationError("context must contain at least one token")   # lm/arithmetic.py:47-48
        return torch.tensor(ids[-1022:], dtype=torch.long)                        # arithmetic.py:90

    def encode_arithmetic_batch(self, bit_lists: Sequence[Sequence[int]], context: Sequence[int], *,
                                quality: Mapping[str, object], eos_stop: bool = False) -> List[List[int]]:
        q = normalise_quality(quality)
        self._reject_rank_policies(q)
        for bits in bit_lists:
            if len(bits) % 8 != 0:
                raise ConfigurationError("bit stream length must be a multiple of 8")
        ctx = self._check_context(context)
        n = len(bit_lists)
        if n == 0:
            return []
        B = self._batch_for(n, decode=False)
        gen = self._generator(B, q)
        msgs = [list(map(int, b)) for b in bit_lists] + [[] for _ in range(B - n)]       # idle streams pad the batch
        # with the sliding KV window (max_len >= 1023) the cover may outgrow the buffer: budget one token per bit
        budget = max((len(b) for b in bit_lists), default=0) + 64 if gen.trunk.ring else None
        tokens = gen.encode(ctx, msgs, max_tokens=budget)
        self._raise_on_status(gen)
        self._last_encode_batch = B
        out = [[int(t) for t in row] for row in tokens[:n]]
        if eos_stop:
            out = [cut_at_eos(row, self.tokenizer) for row in out]
        return out

    def decode_arithmetic_batch(self, token_lists: Sequence[Sequence[int]], context: Sequence[int], *,
                                quality: Mapping[str, object]) -> List[List[int]]:
        q = normalise_quality(quality)
        self._reject_rank_policies(q)
        ctx = self._check_context(context)
        n = len(token_lists)
        if n == 0:
            return []
        B = self._batch_for(n, decode=True)
        gen = self._generator(B, q)
        toks = [list(map(int, t)) for t in token_lists] + [[] for _ in range(B - n)]
        bits = gen.decode(ctx, toks)
        self._raise_on_status(gen)
        return [_trim_to_packet(b) for b in bits[:n]]

    # ------------------------------------------------------------------ cover text -> spans (textio.py:58-63)
    def text_to_spans(self, text: str, seed_text: str, *, quality: Mapping[str, object]) -> List[List[int]]:
        """Parse a cover text back into the token spans ``stego_encode`` produced: re-tokenise, drop the seed, then cut
        the stream where each span's packet closes (and, with ``finish_sent``, its sentence ends), repairing BPE
        re-tokenisation differences with the reference's heuristic (code_base/arithmetic.py:300-342)."""
        from .reveal import SequentialDecoder, split_spans
        from .trunk import StaticGPT2
        q = normalise_quality(quality)
        self._reject_rank_policies(q)
        tok = self.tokenizer
        try:
            ids = [int(t) for t in tok.encode(text, add_special_tokens=False)]
            seed = [int(t) for t in tok.encode(seed_text, add_special_tokens=False)] if seed_text else []
        except TypeError:
            ids = [int(t) for t in tok.encode(text)]
            seed = [int(t) for t in tok.encode(seed_text)] if seed_text else []
        if seed and ids[:len(seed)] == seed:
            ids = ids[len(seed):]
        elif seed_text and text.startswith(seed_text.strip()):
            rest = text[len(seed_text.strip()):]
            try:
                ids = [int(t) for t in tok.encode(rest, add_special_tokens=False)]
            except TypeError:
                ids = [int(t) for t in tok.encode(rest)]
        trunk = StaticGPT2(self.model, self._batch_for(1, decode=True), max_len=self.max_len, device=self.device)
        dec = SequentialDecoder(trunk, tok, precision=int(q["precision"]), temp=float(q["temp"]), topk=int(q["topk"]),
                                device=self.device)
        table = self._sentence_end_table() if q["finish_sent"] else None
        flags = table.cpu() if table is not None else None
        return split_spans(dec, self.encode_seed(seed_text), ids, finish_sent=bool(q["finish_sent"]),
                           is_sentence_end=(lambda t: bool(flags[int(t)])) if flags is not None else (lambda t: True))


def cut_at_eos(tokens: List[int], tokenizer, marker: str = "<eos>") -> List[int]:
    """``code_base/arithmetic.py:206-210``: generation stops with the token that completes ``<eos>`` in the decoded
    cover (text -> bits -> text mode).  The device loop does not look at strings, so the cut is made afterwards."""
    if marker not in tokenizer.decode(tokens):
        return tokens
    lo, hi = 1, len(tokens)                                   # smallest prefix whose text holds the marker
    while lo < hi:
        mid = (lo + hi) // 2
        if marker in tokenizer.decode(tokens[:mid]):
            hi = mid
        else:
            lo = mid + 1
    return tokens[:lo]


class B200RankLM(B200ArithmeticLM):
    """What ``load_lm("gpt2-fa")`` runs in the reference: ``ArithmeticLM`` (lm/arithmetic.py:115-235) on top of the
    rank codec ``encode_with_lm`` / ``decode_with_lm`` (codec/arithmetic.py:122-231), here on the device: KV-cached
    trunk, ``rank_kernel`` per token, no per-token device->host copy.  Decoding needs the side information the
    reference keeps in a FIFO: ``{"history": bits consumed per token, "residual_bits": total bit count as 8 bytes}``
    per chunk (``drain_states`` / ``load_states``, api.py:849-854, :996-1000)."""

    codec = "rank"

    def __init__(self, model, tokenizer, **kw):
        super().__init__(model, tokenizer, **kw)
        from collections import deque
        self._encode_states: List[dict] = []
        self._decode_states = deque()

    def _rank_generator(self, batch: int, q: Mapping[str, object]):
        from .generation import StegoGenerator
        ckw = {}
        V = self.model.config.vocab_size
        # the reference turns top_k into a float and crashes whenever it is below the vocabulary;
        # here it stays an integer and only binds when smaller than V
        if int(q["topk"]) < V:
            ckw["topk"] = int(q["topk"])
        for k in ("top_p", "min_prob", "cap_per_token_bits"):
            if q.get(k) is not None:
                ckw[k] = q[k]
        key = ("rank", batch, float(q["temp"]), tuple(sorted(ckw.items())))
        gen = self._gens.get(key)
        if gen is None:
            gen = StegoGenerator(self.model, batch, max_len=self.max_len, temp=float(q["temp"]), codec="rank", codec_kw=ckw,
                                 device=self.device, use_graph=self.use_graph, trunk_tf32=self.trunk_tf32)
            self._gens = {key: gen}
        return gen

    @staticmethod
    def _msb_bits(bits: Sequence[int]) -> List[int]:
        """API bits are LSB-first per byte (api.py:153-157); the codec reads the payload bytes MSB-first
        (lm/arithmetic.py:16-27 + codec/arithmetic.py:21-60)."""
        payload = bits_to_bytes_lsb(bits)
        return [(byte >> (7 - k)) & 1 for byte in payload for k in range(8)]

    def encode_arithmetic_batch(self, bit_lists, context, *, quality, eos_stop: bool = False):
        q = normalise_quality(quality)
        ctx = self._check_context(context)
        n = len(bit_lists)
        if n == 0:
            return []
        B = self._batch_for(n, decode=False)
        gen = self._rank_generator(B, q)
        msgs = [self._msb_bits(b) for b in bit_lists] + [[] for _ in range(B - n)]
        budget = max((len(b) for b in bit_lists), default=0) + 64 if gen.trunk.ring else None
        tokens = gen.encode(ctx, msgs, max_tokens=budget)
        self._raise_on_status(gen)
        self._last_encode_batch = B
        hist = gen.history_lists()
        for r in range(n):
            state = {"history": tuple(hist[r]), "residual_bits": len(msgs[r]).to_bytes(8, "big")}   # codec/arithmetic.py:165-167
            self._encode_states.append(dict(state))
            self._decode_states.append(dict(state))
        return [[int(t) for t in row] for row in tokens[:n]]

    def decode_arithmetic_batch(self, token_lists, context, *, quality):
        q = normalise_quality(quality)
        ctx = self._check_context(context)
        n = len(token_lists)
        if n == 0:
            return []
        states = []
        for toks in token_lists:
            if not toks:                                      # lm/arithmetic.py:201-203
                if self._decode_states:
                    self._decode_states.popleft()
                states.append(None)
                continue
            if not self._decode_states:
                raise ConfigurationError("decode state unavailable for %s" % type(self).__name__)   # lm/arithmetic.py:204-205
            states.append(self._decode_states.popleft())
        total = []
        for toks, st in zip(token_lists, states):
            if st is None:
                total.append(0)
                continue
            hist = st.get("history", ())
            if len(hist) < len(toks):
                raise ConfigurationError("bit consumption history is required for decoding")      # codec/arithmetic.py:190-191
            tb = int.from_bytes(st.get("residual_bits", b"") or b"", "big")
            total.append(tb if tb else int(sum(hist[: len(toks)])))
        B = self._batch_for(n, decode=True)
        gen = self._rank_generator(B, q)
        toks = [list(map(int, t)) for t in token_lists] + [[] for _ in range(B - n)]
        bits = gen.decode(ctx, toks, total_bits=total + [0] * (B - n))
        self._raise_on_status(gen)
        out = []
        for r in range(n):
            msb = bits[r][: total[r]]
            msb = msb + [0] * ((-len(msb)) % 8)
            payload = bytes(sum(b << (7 - k) for k, b in enumerate(msb[i:i + 8])) for i in range(0, len(msb), 8))
            out.append(bytes_to_bits_lsb(payload))            # lm/arithmetic.py:218-226
        return out

    def text_to_spans(self, text, seed_text, *, quality):
        raise ConfigurationError("the rank codec needs its per-chunk state to decode; text reveal is served by the "
                                 "arithmetic-coder provider")

    def drain_states(self) -> list:                           # lm/arithmetic.py:229-232
        states = [dict(s) for s in self._encode_states]
        self._encode_states.clear()
        return states

    def load_states(self, states) -> None:                    # lm/arithmetic.py:234-235
        from collections import deque
        self._decode_states = deque(dict(s) for s in states)


def random_init_model(name: str = "gpt2", seed: int = 1234):
    """GPT-2 shaped model with random weights (no network in this image): ``gpt2`` = GPT2Config(),
    ``gpt2-fa`` = the HooshvareLab/gpt2-fa shape (42001 tokens).  Seeded like code_base/utils.py:86-88."""
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(seed)
    cfg = GPT2Config() if name == "gpt2" else GPT2Config(vocab_size=42001)
    model = GPT2LMHeadModel(cfg).eval()
    return IdTokenizer(cfg.vocab_size), model


def load_lm(name: str, *, device: Optional[str] = None, max_len: int = 1024, codec: str = "ac",
            batch_size: Optional[int] = None):
    """``mock`` | ``gpt2`` | ``gpt2-fa`` (lm/__init__.py:16-26), plus ``*-random`` offline variants.  ``codec="ac"`` (default)
    gives the arithmetic coder of code_base/arithmetic.py, ``codec="rank"`` the rank codec the reference's
    ``ArithmeticLM`` actually runs (with its ``drain_states`` / ``load_states`` side channel)."""
    n = name.lower()
    if n == "mock":
        return MockLM()
    if codec not in ("ac", "rank"):
        raise ConfigurationError("unknown codec: %s" % codec)
    cls = B200ArithmeticLM if codec == "ac" else B200RankLM
    if n in ("gpt2-random", "gpt2-fa-random"):
        tok, model = random_init_model(n[: -len("-random")])
        return cls(model.to(device or "cuda"), tok, device=device, max_len=max_len, batch_size=batch_size)
    if n in ("gpt2", "gpt2-fa"):
        from transformers import AutoModelForCausalLM, AutoTokenizer
        repo = _MODEL_ALIASES.get(n, n)
        try:
            tok = AutoTokenizer.from_pretrained(repo, local_files_only=True)
            model = AutoModelForCausalLM.from_pretrained(repo, local_files_only=True)
        except Exception as exc:
            raise ConfigurationError("pretrained weights for '%s' are not available offline; use '%s-random' "
                                     "or download the model first" % (repo, n)) from exc
        torch.manual_seed(1234)
        return cls(model.to(device or "cuda"), tok, device=device, max_len=max_len, batch_size=batch_size)
    raise ConfigurationError("unknown language model provider: %s" % name)         # lm/__init__.py:26


__all__ = ["B200ArithmeticLM", "B200RankLM", "MockLM", "IdTokenizer", "load_lm", "normalise_quality", "random_init_model",
           "cut_at_eos"]
