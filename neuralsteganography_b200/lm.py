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

    def __init__(self, model, tokenizer, *, device: Optional[str] = None, max_len: int = 1024, use_graph: bool = True):
        if not torch.cuda.is_available():
            raise ConfigurationError("B200ArithmeticLM needs a CUDA device (there is no CPU fallback)")
        self.model = model.eval()
        self.tokenizer = tokenizer
        self.device = torch.device(device or "cuda")
        self.max_len = int(max_len)
        self.use_graph = bool(use_graph)
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
    def _generator(self, batch: int, q: Mapping[str, object]):
        from .generation import StegoGenerator
        key = (batch, float(q["temp"]), int(q["precision"]), int(q["topk"]), bool(q["finish_sent"]))
        gen = self._gens.get(key)
        if gen is None:
            gen = StegoGenerator(self.model, batch, max_len=self.max_len, precision=int(q["precision"]),
                                 temp=float(q["temp"]), topk=int(q["topk"]), finish_sent=bool(q["finish_sent"]),
                                 sent_end=self._sentence_end_table() if q["finish_sent"] else None,
                                 device=self.device, use_graph=self.use_graph)
            self._gens = {key: gen}            # keep one (KV buffers are large)
        return gen

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
                                quality: Mapping[str, object]) -> List[List[int]]:
        q = normalise_quality(quality)
        for bits in bit_lists:
            if len(bits) % 8 != 0:
                raise ConfigurationError("bit stream length must be a multiple of 8")
        ctx = self._check_context(context)
        gen = self._generator(len(bit_lists), q)
        # with the sliding KV window (max_len >= 1023) the cover may outgrow the buffer: budget one token per bit
        budget = max((len(b) for b in bit_lists), default=0) + 64 if gen.trunk.ring else None
        tokens = gen.encode(ctx, [list(map(int, b)) for b in bit_lists], max_tokens=budget)
        if int((gen.coder.status & 8).sum().item()):
            raise ConfigurationError("cover did not fit %d tokens; raise max_len (>= 1023 slides the window) or shorten "
                                     "the chunk" % gen.coder.token_cap)
        return [[int(t) for t in row] for row in tokens]

    def decode_arithmetic_batch(self, token_lists: Sequence[Sequence[int]], context: Sequence[int], *,
                                quality: Mapping[str, object]) -> List[List[int]]:
        q = normalise_quality(quality)
        ctx = self._check_context(context)
        gen = self._generator(len(token_lists), q)
        bits = gen.decode(ctx, [list(map(int, t)) for t in token_lists])
        return [_trim_to_packet(b) for b in bits]


def random_init_model(name: str = "gpt2", seed: int = 1234):
    """GPT-2 shaped model with random weights (no network in this image): ``gpt2`` = GPT2Config(),
    ``gpt2-fa`` = the HooshvareLab/gpt2-fa shape (42001 tokens).  Seeded like code_base/utils.py:86-88."""
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(seed)
    cfg = GPT2Config() if name == "gpt2" else GPT2Config(vocab_size=42001)
    model = GPT2LMHeadModel(cfg).eval()
    return IdTokenizer(cfg.vocab_size), model


def load_lm(name: str, *, device: Optional[str] = None, max_len: int = 1024):
    """``mock`` | ``gpt2`` | ``gpt2-fa`` (lm/__init__.py:16-26), plus ``*-random`` offline variants."""
    n = name.lower()
    if n == "mock":
        return MockLM()
    if n in ("gpt2-random", "gpt2-fa-random"):
        tok, model = random_init_model(n[: -len("-random")])
        return B200ArithmeticLM(model.to(device or "cuda"), tok, device=device, max_len=max_len)
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
        return B200ArithmeticLM(model.to(device or "cuda"), tok, device=device, max_len=max_len)
    raise ConfigurationError("unknown language model provider: %s" % name)         # lm/__init__.py:26


__all__ = ["B200ArithmeticLM", "MockLM", "IdTokenizer", "load_lm", "normalise_quality", "random_init_model"]
