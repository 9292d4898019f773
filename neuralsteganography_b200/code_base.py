"""The reference's research-script surface on top of the device coder.

``encode_arithmetic`` / ``decode_arithmetic`` (code_base/arithmetic.py:78-88, :220-229), ``encode_huffman`` /
``decode_huffman`` (code_base/huffman_baseline.py:7, :73), ``get_bins`` / ``encode_block`` / ``decode_block``
(code_base/block_baseline.py:9, :26, :99), ``sample`` (code_base/sample.py:6) and the ``run_single`` driver
(code_base/run_single.py:11-95) keep their argument order and return tuples.  ``model`` is a HuggingFace GPT-2 shaped
module; the trunk runs through :class:`~neuralsteganography_b200.trunk.StaticGPT2` on ``device``, every token choice
is made by the CUDA kernels.  Single-stream, host-paced paths: they exist for drop-in compatibility, the batched
provider is the throughput path.
"""

from __future__ import annotations

import heapq
import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from .codecs import CodecStreams, get_bins  # noqa: F401  (get_bins is part of the mirrored surface)
from .generation import StegoGenerator
from .lm import cut_at_eos
from .trunk import StaticGPT2

LN2 = 0.69315          # the reference's constant (code_base/utils.py:33,38)


def encode_arithmetic(model, enc, message: Sequence[int], context: Sequence[int], finish_sent: bool = False,
                      device: str = "cuda", temp: float = 1.0, precision: int = 16, topk: int = 50000,
                      max_len: int = 1024):
    """-> (tokens, avg_NLL, avg_KL, words_per_bit, avg_Hq) like the reference (arithmetic.py:212-217).
    The statistics are accumulated on the device, one [1,3] read per step (this entry point is the
    single-stream research path; the batched provider does not compute them).  Generation ends with the token that
    completes ``<eos>`` in the decoded cover (:206-210, text -> bits -> text mode)."""
    gen = StegoGenerator(model, 1, max_len=max_len, precision=precision, temp=temp, topk=topk,
                         finish_sent=finish_sent, device=device, use_graph=False, collect_stats=True)
    ctx = torch.tensor(list(context)[-1022:], dtype=torch.long)

    def eos_poll(coder, t):                               # every few steps: has the cover spelled "<eos>" yet?
        toks = coder.token_lists()[0]
        return hasattr(enc, "decode") and "<eos>" in enc.decode(toks)

    budget = len(message) + 64 if gen.trunk.ring else None
    tokens = gen.encode(ctx, [list(map(int, message))], max_tokens=budget, poll_every=8,
                        on_poll=eos_poll if hasattr(enc, "decode") else None)[0]
    if hasattr(enc, "decode"):
        tokens = cut_at_eos(tokens, enc)
    used = int(gen.coder.cursor[0].item())
    n = max(1, gen.stats_steps)
    nan = float("nan")
    lp, kl, hq = [float(x) for x in gen.stats_sum[0].tolist()]
    return tokens, -lp / n, kl / n, (gen.stats_steps / used if used else nan), hq / n


def decode_arithmetic(model, enc, text, context: Sequence[int], device: str = "cuda", temp: float = 1.0,
                      precision: int = 16, topk: int = 50000, max_len: int = 1024) -> List[int]:
    """``text`` may be the cover string (re-tokenised with ``enc``, with the reference's BPE repair :234-242, :300-342)
    or the token ids themselves."""
    from .reveal import SequentialDecoder
    inp = enc.encode(text) if isinstance(text, str) else [int(t) for t in text]
    trunk = StaticGPT2(model, 1, max_len=max(max_len, min(1024, len(context) + 2 * len(inp) + 80)), device=device)
    # encode_arithmetic collects the reference's statistics, which only the exact kernel computes: decode with the same
    # kernel, so that both directions share one summation tree (the throughput kernel agrees with it to ~1e-7 per row,
    # not by construction)
    dec = SequentialDecoder(trunk, enc, precision=precision, temp=temp, topk=topk, device=device, force_exact=True)
    bits, _tokens, _used = dec.run(list(context), inp)
    return bits


# ---------------------------------------------------------------------------------------------- baselines
def _huffman_code_lengths(freqs: List[float]) -> List[int]:
    """Code lengths of the reference's tree (huffman.py:43-76: heapq of nodes ordered by frequency only)."""

    class _Node:
        __slots__ = ("token", "freq", "left", "right")

        def __init__(self, token, freq):
            self.token, self.freq, self.left, self.right = token, freq, None, None

        def __lt__(self, other):
            return self.freq < other.freq

    heap: list = []
    for idx, f in enumerate(freqs):
        heapq.heappush(heap, _Node(idx, f))
    while len(heap) > 1:
        a, b = heapq.heappop(heap), heapq.heappop(heap)
        m = _Node(None, a.freq + b.freq)
        m.left, m.right = a, b
        heapq.heappush(heap, m)
    lengths = [0] * len(freqs)
    stack = [(heap[0], 0)]
    while stack:
        node, d = stack.pop()
        if node.token is not None:
            lengths[node.token] = d
        else:
            stack.append((node.left, d + 1))
            stack.append((node.right, d + 1))
    return lengths


def _kl_bits(q: torch.Tensor, logq: torch.Tensor, logp: torch.Tensor) -> float:
    res = q * (logq - logp) / LN2                                   # utils.py:32-35
    res[q == 0] = 0
    return float(res.sum().item())


def _masked(logits: torch.Tensor) -> torch.Tensor:
    x = logits.clone()
    x[-1] = -1e10                                                   # huffman_baseline.py:26-27
    if x.numel() > 628:
        x[628] = -1e10
    return x


def _lut_from(words2bin, vocab: int, device) -> Optional[torch.Tensor]:
    if words2bin is None:
        return None
    lut = torch.full((vocab,), -1, dtype=torch.int32)
    if isinstance(words2bin, dict):
        for w, b in words2bin.items():
            lut[int(w)] = int(b)
    else:
        lut[: len(words2bin)] = torch.as_tensor(words2bin, dtype=torch.int32)
    return lut.to(device)


def _baseline_encode(kind: str, model, enc, message, context, param: int, device: str, finish_sent: bool,
                     words2bin=None, max_len: int = 1024):
    trunk = StaticGPT2(model, 1, max_len=max_len, device=device)
    ctx = torch.tensor(list(context)[-1022:], dtype=torch.long, device=device)[None]
    st = CodecStreams(kind, 1, trunk.vocab, param=param, device=device, token_cap=max_len)
    lut = _lut_from(words2bin, trunk.vocab, device)
    if lut is not None:
        st.lut = lut                                                # caller's bins (block_baseline.py:26 signature)
    st.set_messages([list(map(int, message))])
    logits = trunk.prefill(ctx)
    room = max_len - ctx.shape[1] - 1
    total_lp, total_kl, nstat = 0.0, 0.0, 0
    tail: List[int] = []
    steps = 0
    while steps < room:
        row = _masked(logits[0])
        if st.all_done():                                           # message consumed
            if not finish_sent:
                break
            top = int(torch.argmax(row).item())                     # rank-0 token (:37-39 / :52-54)
            tail.append(top)
            steps += 1
            text = enc.decode([top]) if hasattr(enc, "decode") else "."
            if "." in text or "!" in text or "?" in text:           # utils.py:55-57
                break
            logits = trunk.step(torch.tensor([top], device=device))
            continue
        st.encode_step(logits)
        tok = int(st.tokens[0, int(st.ntok[0].item()) - 1].item())
        logp = F.log_softmax(row, dim=-1)
        if kind == "huffman":                                       # huffman_baseline.py:30-34, :57-63
            n = min(1 << param, row.numel())
            vals, idx = torch.sort(row, descending=True, stable=True)
            lp_top = F.log_softmax(vals, dim=-1)[:n]
            lengths = _huffman_code_lengths(torch.exp(lp_top).cpu().numpy().tolist())
            logq = torch.tensor([-float(l) for l in lengths], device=device) * LN2
            total_kl += _kl_bits(torch.exp(logq), logq, lp_top)
        else:                                                       # block_baseline.py:55-71: one token per bin, 2^-b each
            table = st.lut.long()
            nb = 1 << param
            per_bin = torch.where(table[None, :] == torch.arange(nb, device=device)[:, None], row[None, :],
                                  torch.full((), -float("inf"), device=device))
            best, arg = per_bin.max(dim=1)
            ok = torch.isfinite(best)
            total_kl += float(((2.0 ** -param) * ((-param * LN2) - logp[arg[ok]]) / LN2).sum().item())
        total_lp += float(logp[tok].item())
        nstat += 1
        steps += 1
        logits = trunk.step(torch.tensor([tok], device=device))
    used = int(st.cursor[0].item())
    if used < len(message):
        raise ValueError("the cover ran out of room (%d tokens) before the message was consumed: %d of %d bits coded; raise "
                         "max_len" % (room, used, len(message)))
    tokens = st.token_lists()[0] + tail
    nan = float("nan")
    n = max(1, nstat)
    return tokens, -total_lp / n, total_kl / n, (nstat / used if used else nan)


def _baseline_decode(kind: str, model, enc, text, context, param: int, device: str, words2bin=None,
                     max_len: int = 1024) -> List[int]:
    inp = enc.encode(text) if isinstance(text, str) else [int(t) for t in text]
    out: List[int] = []
    i = 0
    while i < len(inp):                       # 628 -> 198 198 (huffman_baseline.py:78-85, block_baseline.py:104-111)
        if inp[i] == 628:
            inp[i] = 198
            inp[i + 1:i + 1] = [198]
            i += 2
        else:
            i += 1
    ctx = torch.tensor(list(context)[-1022:], dtype=torch.long, device=device)[None]
    need = ctx.shape[1] + len(inp) + 1
    if need > max_len:
        raise ValueError("decoding %d tokens after a %d-token context needs max_len >= %d" % (len(inp), ctx.shape[1], need))
    trunk = StaticGPT2(model, 1, max_len=max_len, device=device)
    st = CodecStreams(kind, 1, trunk.vocab, param=param, device=device, token_cap=max(1, len(inp)))
    lut = _lut_from(words2bin, trunk.vocab, device)
    if lut is not None:
        st.lut = lut
    st.set_tokens([inp])
    logits = trunk.prefill(ctx)
    for t in range(len(inp)):
        st.decode_step(logits)
        logits = trunk.step(torch.tensor([inp[t]], device=device))
    return st.bit_lists()[0]


def encode_huffman(model, enc, message, context, bits_per_word, finish_sent=False, device="cuda"):
    """-> (tokens, avg_NLL, avg_KL, words_per_bit) (huffman_baseline.py:66-71)."""
    return _baseline_encode("huffman", model, enc, message, context, bits_per_word, device, finish_sent)


def decode_huffman(model, enc, text, context, bits_per_word, device="cuda"):
    return _baseline_decode("huffman", model, enc, text, context, bits_per_word, device)


def encode_block(model, enc, message, context, block_size, bin2words=None, words2bin=None, finish_sent=False, device="cuda"):
    """-> (tokens, avg_NLL, avg_KL, words_per_bit) (block_baseline.py:92-97); ``words2bin`` (dict or table) overrides
    the bins of :func:`get_bins`."""
    return _baseline_encode("bins", model, enc, message, context, block_size, device, finish_sent, words2bin=words2bin)


def decode_block(model, enc, text, context, block_size, bin2words=None, words2bin=None, device="cuda"):
    return _baseline_decode("bins", model, enc, text, context, block_size, device, words2bin=words2bin)


def sample(model, enc, length, context, temperature=1.0, device="cuda", topk=-1, max_len: int = 1024, seed: Optional[int] = None):
    """Plain temperature / top-k sampling with the same statistics (code_base/sample.py:6-55): no message is embedded --
    the quality baseline of the comparison tables.  -> (tokens, avg_NLL, avg_KL, avg_Hq)."""
    assert length > 0
    ctx = torch.tensor(list(context)[-1022:], dtype=torch.long, device=device)[None]
    if ctx.shape[1] + length + 1 > max_len and max_len < 1023:
        raise RuntimeError("sample: context + length exceed the KV buffer")          # sample.py:23-24
    trunk = StaticGPT2(model, 1, max_len=max_len, device=device)
    g = torch.Generator(device=device)
    if seed is not None:
        g.manual_seed(seed)
    logits = trunk.prefill(ctx)
    out: List[int] = []
    tlp = tkl = tent = 0.0
    for _ in range(length):
        vals, idx = torch.sort(_masked(logits[0]), descending=True, stable=True)
        base = F.log_softmax(vals, dim=-1)
        v = vals[:topk] if topk > 0 else vals
        lp = F.log_softmax(v / temperature, dim=-1)
        p = torch.exp(lp)
        tkl += _kl_bits(p, lp, base[: p.numel()] if topk > 0 else base[:-1])          # sample.py:40 (base[:topk])
        sel = int(torch.multinomial(p, 1, generator=g).item())
        tlp += float(base[sel].item())
        e = p * lp / LN2
        e[p == 0] = 0
        tent += -float(e.sum().item())
        tok = int(idx[sel].item())
        out.append(tok)
        logits = trunk.step(torch.tensor([tok], device=device))
    return out, -tlp / length, tkl / length, tent / length


def run_single(model, enc, message_str: str = "This is a very secret message!", *, context_tokens: Sequence[int],
               mode: str = "arithmetic", block_size: int = 3, temp: float = 0.9, precision: int = 26, sample_tokens: int = 100,
               topk: int = 300, finish_sent: bool = False, device: str = "cuda", text_precision: int = 40,
               text_topk: int = 60000) -> Dict[str, object]:
    """The reference's driver (code_base/run_single.py:11-95): text -> uniform bits by arithmetic DEcoding the message
    itself at precision 40 / topk 60000 (:53-54), bits -> cover, cover -> bits, bits -> text by arithmetic ENcoding
    until ``<eos>`` (:93-94)."""
    if mode not in ("arithmetic", "huffman", "bins", "sample"):
        raise NotImplementedError(mode)
    message_ctx = enc.encode("<|endoftext|>")
    message = decode_arithmetic(model, enc, message_str + "<eos>", message_ctx, device=device, precision=text_precision,
                                topk=text_topk)
    hq = 0.0
    if mode == "arithmetic":
        out, nll, kl, wpb, hq = encode_arithmetic(model, enc, message, context_tokens, temp=temp, finish_sent=finish_sent,
                                                  precision=precision, topk=topk, device=device)
    elif mode == "huffman":
        out, nll, kl, wpb = encode_huffman(model, enc, message, context_tokens, block_size, finish_sent=finish_sent, device=device)
    elif mode == "bins":
        out, nll, kl, wpb = encode_block(model, enc, message, context_tokens, block_size, finish_sent=finish_sent, device=device)
    else:
        out, nll, kl, hq = sample(model, enc, sample_tokens, context_tokens, temperature=temp, topk=topk, device=device)
        wpb = 1
    text = enc.decode(out)
    res: Dict[str, object] = {"message_bits": message, "cover_tokens": out, "cover_text": text, "ppl": math.exp(nll), "kl": kl,
                              "words_per_bit": wpb, "entropy": hq / LN2}
    if mode != "sample":
        if mode == "arithmetic":
            rec = decode_arithmetic(model, enc, text, context_tokens, temp=temp, precision=precision, topk=topk, device=device)
        elif mode == "huffman":
            rec = decode_huffman(model, enc, text, context_tokens, block_size, device=device)
        else:
            rec = decode_block(model, enc, text, context_tokens, block_size, device=device)
        reconst = encode_arithmetic(model, enc, rec, message_ctx, precision=text_precision, topk=text_topk, device=device)
        res["recovered_bits"] = rec
        res["reconstructed_text"] = enc.decode(reconst[0])
    return res
