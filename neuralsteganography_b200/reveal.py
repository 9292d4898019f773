"""Cover text -> token spans -> bits: the reveal side the reference leaves unfinished.

* ``decode_with_repair``  the decode loop of ``code_base/arithmetic.py:255-371`` INCLUDING its BPE-repair heuristic
  (``:234-242`` the 628 -> 198,198 pre-pass, ``:300-342`` "a more likely prefix / longer token"): re-tokenising a cover
  text does not always give back the tokens that were generated.  The coder step stays the CUDA kernel; when it
  flags an observed token outside the kept set (``NS_ST_OUT_OF_RANGE``) the host applies the reference's string
  heuristic to the ranked candidates, patches the token list, restores the stream's state and re-steps.
* ``text_to_spans``       ``src/neuralstego/codec/textio.py:58-63`` (``NotImplementedError`` in the reference): the
  cover is one token stream ``seed + span_0 + span_1 + ...`` (``spans_to_text`` :36-55); spans are delimited in band --
  a span's packet is complete when its JSON object closes (``codec/packet.py:97-106``), after which ``finish_sent``
  covers run on to the first sentence-ending token (``code_base/arithmetic.py:114,135-137``).

Single-stream, host-paced paths (one small device->host read per token): they serve ``main.py``-style reveal of one
text, not the batched throughput path.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _native as N
from .coder import ArithmeticStreams, unpack_bits
from .exceptions import ConfigurationError

NEWLINE2, NEWLINE1, BYTE_NL = 628, 198, 128          # code_base/arithmetic.py:125,236-241,307


def prepass_628(inp: List[int]) -> List[int]:
    """``code_base/arithmetic.py:234-242``: the tokenizer merges two newlines into token 628, which the coder forbids."""
    out: List[int] = []
    for t in inp:
        if t == NEWLINE2:
            out += [NEWLINE1, NEWLINE1]
        else:
            out.append(int(t))
    return out


def bpe_repair(inp: List[int], i: int, ranked: Sequence[int], enc) -> Optional[int]:
    """The reference's heuristic (``code_base/arithmetic.py:300-342``) for an observed token ``inp[i]`` that is not among
    the kept candidates ``ranked`` (in coder order).  Patches ``inp`` in place and returns the rank to code, or ``None``
    when nothing fits (the reference then codes rank 0)."""
    true_text = enc.decode([inp[i]])
    for rank_idx, cand in enumerate(ranked):
        cand = int(cand)
        prop = enc.decode([cand])
        if inp[i] == BYTE_NL and cand == NEWLINE1:                       # :307-310
            inp[i] = cand
            return rank_idx
        if len(prop) <= len(true_text) and prop == true_text[:len(prop)]:   # a more likely prefix token (:313-319)
            suffix_tokens = [int(t) for t in enc.encode(true_text[len(prop):])]
            inp[i] = cand
            inp[i + 1:i + 1] = suffix_tokens
            return rank_idx
        if len(prop) > len(true_text) and true_text == prop[:len(true_text)]:   # a more likely longer token (:322-338)
            whole, extra = true_text, 1
            while len(whole) < len(prop) and i + extra < len(inp):
                whole += enc.decode([inp[i + extra]])
                extra += 1
            if prop == whole[:len(prop)]:
                inp[i] = cand
                del inp[i + 1:i + extra]
                if len(whole) > len(prop):
                    inp[i + 1:i + 1] = [int(t) for t in enc.encode(whole[len(prop):])]
                return rank_idx
    return None


def _packet_bytes_complete(data: bytes) -> int:
    """Length of the JSON packet at the head of ``data`` once its object has closed, else 0."""
    if data[:1] != b"{":
        return 0
    depth, in_str, esc = 0, False, False
    for k, ch in enumerate(data):
        if in_str:
            if esc:
                esc = False
            elif ch == 0x5C:
                esc = True
            elif ch == 0x22:
                in_str = False
        elif ch == 0x22:
            in_str = True
        elif ch == 0x7B:
            depth += 1
        elif ch == 0x7D:
            depth -= 1
            if depth == 0:
                return k + 1
    return 0


def _lsb_bytes(bits: Sequence[int]) -> bytes:
    out = bytearray()
    for k in range(0, len(bits) - len(bits) % 8, 8):
        v = 0
        for off in range(8):
            v |= (bits[k + off] & 1) << off
        out.append(v)
    return bytes(out)


class SequentialDecoder:
    """One stream, one token at a time: trunk step + CUDA coder step, with the reference's BPE repair in between."""

    def __init__(self, trunk, tokenizer, *, precision: int, temp: float, topk: int, device="cuda", force_exact: bool = False):
        # the trunk may be wider than one stream: the stream is then replicated over its rows so that the GEMM shapes
        # -- and with them the fp32 summation order of the logits -- are those of the batch that encoded the cover
        self.trunk, self.enc = trunk, tokenizer
        self.V = trunk.vocab
        self.precision, self.temp, self.topk = int(precision), float(temp), int(topk)
        self.device = torch.device(device)
        self.force_exact = bool(force_exact)            # pair with an encoder that ran the exact kernel (statistics)
        self.repairs = 0
        self.unrepaired = 0

    def _ranked_candidates(self, logits: torch.Tensor, lo: int, hi: int, mask_ids) -> List[int]:
        """The kept tokens in coder order for the host heuristic (device sort; code_base/arithmetic.py:127-142)."""
        x = logits.clone()
        for m in mask_ids:
            if 0 <= m < self.V:
                x[m] = -float("inf")
        vals, idx = torch.sort(x, descending=True, stable=True)          # equal logits: lower id first
        p = torch.softmax(vals.double() / self.temp, dim=0)
        below = (p < 1.0 / float(hi - lo)).nonzero()
        k = int(below[0].item()) if len(below) else self.V
        k = min(max(2, k), self.topk)
        return idx[:k].tolist()

    def run(self, context: Sequence[int], tokens: Sequence[int], *, stop: Optional[Callable[[List[int], int, List[int]], bool]] = None,
            flush_last: bool = True) -> Tuple[List[int], List[int], int]:
        """Decode ``tokens`` after ``context``.  ``stop(bits, i, inp)`` is asked after every token and ends the stream
        when it returns True.  Returns ``(bits, repaired token list, tokens consumed)``."""
        inp = prepass_628([int(t) for t in tokens])
        cap = len(inp) + len(inp) // 2 + 64
        st = ArithmeticStreams(1, self.V, precision=self.precision, temp=self.temp, topk=self.topk, token_cap=cap,
                               device=self.device, force_exact=self.force_exact)
        st.set_tokens([inp])
        if not flush_last:
            st.ntok_total.fill_(cap + 1)                    # no token is "the last one" (arithmetic.py:356): spans end in band
        ctx = torch.tensor([int(t) for t in context][-1022:], dtype=torch.long, device=self.device)[None]
        ctx = ctx.expand(self.trunk.B, -1).contiguous()
        self.trunk.reset()
        logits = self.trunk.prefill(ctx)[:1]
        i = 0
        bits: List[int] = []
        while i < len(inp):
            snap = (st.lo.clone(), st.hi.clone(), st.out_len.clone(), st.ntok.clone(), st.phase.clone(), st.out_bits.clone())
            st.status.zero_()
            st.decode_step(logits)
            if int(st.status[0].item()) & N.ST_OUT_OF_RANGE:
                ranked = self._ranked_candidates(logits[0], int(snap[0][0].item()), int(snap[1][0].item()), st.mask_ids)
                rank = bpe_repair(inp, i, ranked, self.enc)
                if rank is not None:
                    if len(inp) > cap:
                        raise ConfigurationError("cover text needs more token slots than allocated for its repair")
                    self.repairs += 1
                    st.lo.copy_(snap[0]); st.hi.copy_(snap[1]); st.out_len.copy_(snap[2]); st.ntok.copy_(snap[3])
                    st.phase.copy_(snap[4]); st.out_bits.copy_(snap[5])
                    row = torch.full((cap,), -1, dtype=torch.int32)
                    row[: len(inp)] = torch.tensor(inp, dtype=torch.int32)
                    st.tokens[0].copy_(row.to(self.device))
                    if flush_last:
                        st.ntok_total.fill_(len(inp))
                    st.status.zero_()
                    st.decode_step(logits)
                else:
                    self.unrepaired += 1                    # "Unable to fix BPE error": coded as rank 0 (:340-342)
            i += 1
            if stop is not None or i == len(inp):
                words = st.out_bits.cpu().numpy().view("uint32")
                bits = unpack_bits(words, st.out_len.cpu().numpy())[0]
                if stop is not None and stop(bits, i, inp):
                    break
            if i < len(inp):
                # same live KV prefix as the generation loop used at this position (generation.py _run)
                logits = self.trunk.step(torch.full((self.trunk.B,), inp[i - 1], dtype=torch.long, device=self.device),
                                         self.trunk.kv_bucket(int(ctx.shape[1]) + i))[:1]
        return bits, inp, i


def split_spans(decoder: SequentialDecoder, context: Sequence[int], tokens: Sequence[int], *, finish_sent: bool,
                is_sentence_end: Callable[[int], bool], max_spans: int = 1 << 20) -> List[List[int]]:
    """Cut the cover's token stream into the spans ``stego_encode`` produced (each coded after the same context)."""
    rest = [int(t) for t in tokens]
    spans: List[List[int]] = []
    while rest and len(spans) < max_spans:
        state = {"end": None}

        def stop(bits, i, inp, state=state):
            if state["end"] is None:
                if _packet_bytes_complete(_lsb_bytes(bits)):
                    state["end"] = i                            # the packet closed with token i-1
                    if not finish_sent:
                        return True
                    return False                                # the tail starts with the next token (:135-137)
                return False
            return is_sentence_end(inp[i - 1])                  # tail: stop after the first sentence-ending token

        _bits, inp, used = decoder.run(context, rest, stop=stop, flush_last=False)
        if state["end"] is None:
            raise ConfigurationError("cover text ends inside a packet (span %d is incomplete)" % len(spans))
        spans.append(inp[:used])
        rest = inp[used:]
    return spans


__all__ = ["SequentialDecoder", "split_spans", "bpe_repair", "prepass_628"]
