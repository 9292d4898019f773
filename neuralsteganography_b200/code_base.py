"""The reference's research-script signatures on top of the device coder.

``encode_arithmetic`` / ``decode_arithmetic`` (code_base/arithmetic.py:78-88, :220-229),
``encode_huffman`` / ``decode_huffman`` (code_base/huffman_baseline.py:7, :73) and
``get_bins`` / ``encode_block`` / ``decode_block`` (code_base/block_baseline.py:9, :26, :99) keep
their argument order and return tuples.  ``model`` is a HuggingFace GPT-2 shaped module; the trunk
runs through :class:`~neuralsteganography_b200.trunk.StaticGPT2` on ``device``.
"""

from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .codecs import CodecStreams, get_bins  # noqa: F401  (get_bins is part of the mirrored surface)
from .generation import StegoGenerator
from .trunk import StaticGPT2


def encode_arithmetic(model, enc, message: Sequence[int], context: Sequence[int], finish_sent: bool = False,
                      device: str = "cuda", temp: float = 1.0, precision: int = 16, topk: int = 50000,
                      max_len: int = 1024):
    """-> (tokens, avg_NLL, avg_KL, words_per_bit, avg_Hq) like the reference (arithmetic.py:212-217).
    The statistics are accumulated on the device, one [1,3] read per step (this entry point is the
    single-stream research path; the batched provider does not compute them)."""
    gen = StegoGenerator(model, 1, max_len=max_len, precision=precision, temp=temp, topk=topk,
                         finish_sent=finish_sent, device=device, use_graph=False, collect_stats=True)
    ctx = torch.tensor(list(context)[-1022:], dtype=torch.long)
    tokens = gen.encode(ctx, [list(map(int, message))])[0]
    used = int(gen.coder.cursor[0].item())
    n = max(1, gen.stats_steps)
    nan = float("nan")
    lp, kl, hq = [float(x) for x in gen.stats_sum[0].tolist()]
    return tokens, -lp / n, kl / n, (gen.stats_steps / used if used else nan), hq / n


def decode_arithmetic(model, enc, text, context: Sequence[int], device: str = "cuda", temp: float = 1.0,
                      precision: int = 16, topk: int = 50000, max_len: int = 1024) -> List[int]:
    """``text`` may be the cover string (re-tokenised with ``enc``) or the token ids themselves."""
    inp = enc.encode(text) if isinstance(text, str) else [int(t) for t in text]
    i = 0
    while i < len(inp):                       # 628 -> 198 198 repair of the reference, arithmetic.py:234-242
        if inp[i] == 628:
            inp[i] = 198
            inp[i + 1:i + 1] = [198]
            i += 2
        else:
            i += 1
    gen = StegoGenerator(model, 1, max_len=max_len, precision=precision, temp=temp, topk=topk, device=device)
    ctx = torch.tensor(list(context)[-1022:], dtype=torch.long)
    return gen.decode(ctx, [inp])[0]


def _baseline_loop(kind: str, model, message, context, param: int, device: str, decode_tokens=None, max_len: int = 1024):
    trunk = StaticGPT2(model, 1, max_len=max_len, device=device)
    ctx = torch.tensor(list(context)[-1022:], dtype=torch.long, device=device)[None]
    st = CodecStreams(kind, 1, trunk.vocab, param=param, device=device, token_cap=max_len)
    logits = trunk.prefill(ctx)
    if decode_tokens is None:
        st.set_messages([list(map(int, message))])
        for _ in range(max_len - ctx.shape[1] - 1):
            st.encode_step(logits)
            if st.all_done():
                break
            logits = trunk.step(st.tokens[:, int(st.ntok[0].item()) - 1].long())
        return st.token_lists()[0], int(st.cursor[0].item())
    st.set_tokens([decode_tokens])
    for t in range(len(decode_tokens)):
        st.decode_step(logits)
        logits = trunk.step(torch.tensor([decode_tokens[t]], device=device))
    return st.bit_lists()[0]


def encode_huffman(model, enc, message, context, bits_per_word, finish_sent=False, device="cuda"):
    tokens, used = _baseline_loop("huffman", model, message, context, bits_per_word, device)
    nan = float("nan")
    return tokens, nan, nan, (len(tokens) / used if used else nan)


def decode_huffman(model, enc, text, context, bits_per_word, device="cuda"):
    inp = enc.encode(text) if isinstance(text, str) else [int(t) for t in text]
    return _baseline_loop("huffman", model, None, context, bits_per_word, device, decode_tokens=inp)


def encode_block(model, enc, message, context, block_size, bin2words=None, words2bin=None, finish_sent=False, device="cuda"):
    tokens, used = _baseline_loop("bins", model, message, context, block_size, device)
    nan = float("nan")
    return tokens, nan, nan, (len(tokens) / used if used else nan)


def decode_block(model, enc, text, context, block_size, bin2words=None, words2bin=None, device="cuda"):
    inp = enc.encode(text) if isinstance(text, str) else [int(t) for t in text]
    return _baseline_loop("bins", model, None, context, block_size, device, decode_tokens=inp)
