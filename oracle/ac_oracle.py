"""Numpy restatement of the reference's finite-precision arithmetic coder (A).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites
the reference lines it follows; paths are relative to ``/root/reference``.

The coder works on one stream.  Each step takes one fp32 logits row ``[V]``
(what ``model(...).logits[0, -1, :]`` would be) and the coder state
``(lo, hi)``; it never touches a language model, so the same rows can be fed
to the live reference (``oracle/ref_harness.py``), to this restatement and to
the CUDA kernel.

Tie-break: the reference sorts with ``torch.sort`` (code_base/arithmetic.py:127),
whose order among equal logits is implementation-defined.  The oracle (and the
kernel) fix "equal logits: lower token id first".
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

ENC_MASK = np.float32(-1e20)   # code_base/arithmetic.py:124-125
DEC_MASK = np.float32(-1e10)   # code_base/arithmetic.py:265-266
NEWLINE2_ID = 628              # code_base/arithmetic.py:125


# ----------------------------------------------------------------------------
# bit helpers -- code_base/utils.py:43-64
# ----------------------------------------------------------------------------
def bits2int(bits: Sequence[int]) -> int:
    """LSB-first bits -> int (code_base/utils.py:43-47)."""
    res = 0
    for i, bit in enumerate(bits):
        res += int(bit) << i
    return res


def int2bits(value: int, num_bits: int) -> List[int]:
    """int -> LSB-first list of ``num_bits`` bits (code_base/utils.py:49-53)."""
    if num_bits == 0:
        return []
    text = format(int(value), "0%db" % num_bits)
    return [int(ch) for ch in reversed(text)]


def num_same_from_beg(bits1: Sequence[int], bits2: Sequence[int]) -> int:
    """Common-prefix length, capped at ``len-1`` (code_base/utils.py:59-64).

    The reference's loop never breaks on identical lists, so it returns the
    last loop index ``len-1`` for them; that quirk is part of the format.
    """
    assert len(bits1) == len(bits2)
    i = 0
    for i in range(len(bits1)):
        if bits1[i] != bits2[i]:
            break
    return i


# ----------------------------------------------------------------------------
# distribution -- code_base/arithmetic.py:124-158 (encode), :265-296 (decode)
# ----------------------------------------------------------------------------
def mask_row(row: np.ndarray, mask_value: np.float32) -> np.ndarray:
    """Forbid the last token and token 628 (code_base/arithmetic.py:124-125)."""
    out = np.array(row, dtype=np.float32, copy=True)
    out[-1] = mask_value
    if out.shape[0] > NEWLINE2_ID:
        out[NEWLINE2_ID] = mask_value
    return out


def sort_desc(row: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Descending sort, equal values by ascending id (arithmetic.py:127)."""
    order = np.argsort(-row, kind="stable")
    return row[order], order


def softmax_f64(sorted_logits: np.ndarray, temp: float) -> np.ndarray:
    """fp64 softmax of ``double(s)/temp`` (arithmetic.py:128-130).

    torch's CPU kernel computes ``exp(x - max) * (1 / sum)``; the summation
    order is torch-internal (SIMD-width dependent), so agreement with the
    reference is to within a few ulp -- see DESIGN.md "mismatch rate".
    """
    x = sorted_logits.astype(np.float64) / float(temp)
    e = np.exp(x - x.max())
    return e * (1.0 / e.sum())


def select_cutoff_k(probs: np.ndarray, threshold: float, topk: int) -> int:
    """``min(max(2, first idx with p < thr else len), topk)`` (arithmetic.py:51-75)."""
    below = np.nonzero(probs < threshold)[0]
    candidate = int(below[0]) if below.size else int(probs.shape[0])
    return min(max(2, candidate), int(topk))


def integer_cdf(probs: np.ndarray, cur_range: int, topk: int) -> Tuple[np.ndarray, int]:
    """Integer CDF over the kept bins, relative to ``lo`` (arithmetic.py:140-158).

    Returns ``(cum, k)``: ``cum[j]`` is the exclusive top of bin ``j`` minus
    ``lo``; ``k = len(cum)`` after the overfill truncation (decode, :290).
    """
    thr = 1.0 / cur_range                                   # :141
    k = select_cutoff_k(probs, thr, topk)                   # :142
    kept = probs[:k]
    scaled = kept / kept.sum() * cur_range                  # :146
    q = np.rint(scaled).astype(np.int64)                    # :149 (half-to-even)
    cum = np.cumsum(q)                                      # :150
    over = np.nonzero(cum > cur_range)[0]                   # :153
    if over.size:
        cum = cum[: over[0]]                                # :155
        k = int(over[0])                                    # :290
    cum = cum + (cur_range - int(cum[-1]))                  # :158
    return cum, k


def interval_update(new_bottom: int, new_top: int, precision: int) -> Tuple[int, int, int, List[int], List[int]]:
    """Shared-prefix emission and rescale (arithmetic.py:179-190).

    Returns ``(n, lo', hi', bottom_bits_msb_first, top_bits_msb_first)``.
    """
    bottom_bits = list(reversed(int2bits(new_bottom, precision)))       # :179
    top_bits = list(reversed(int2bits(new_top - 1, precision)))         # :180
    n = num_same_from_beg(bottom_bits, top_bits)                        # :183
    nb_bits = bottom_bits[n:] + [0] * n                                 # :186
    nt_bits = top_bits[n:] + [1] * n                                    # :187
    lo = bits2int(reversed(nb_bits))                                    # :189
    hi = bits2int(reversed(nt_bits)) + 1                                # :190
    return n, lo, hi, bottom_bits, top_bits


@dataclass
class StepTrace:
    """Per-step record used by the parity tests."""

    token: int
    selection: int
    k: int
    nbits: int
    lo: int
    hi: int
    new_bottom: int
    new_top: int


@dataclass
class EncodeResult:
    tokens: List[int]
    bits_consumed: int
    trace: List[StepTrace] = field(default_factory=list)
    avg_nll: float = float("nan")
    avg_kl: float = float("nan")
    words_per_bit: float = float("nan")
    avg_hq: float = float("nan")


def _stats(sorted_logits: np.ndarray, temp: float, probs_temp: np.ndarray,
           cum: np.ndarray, selection: int) -> Tuple[float, float, float]:
    """log p(sel), KL(q||p) bits, H(p_temp) (arithmetic.py:131-132,192-198; utils.py:32-40)."""
    s64 = sorted_logits.astype(np.float64)

    def log_softmax(x: np.ndarray) -> np.ndarray:
        z = x - x.max()
        return z - math.log(np.exp(z).sum())

    log_probs_temp = log_softmax(s64 / float(temp))
    log_probs = log_softmax(s64)
    widths = cum.astype(np.float64).copy()
    widths[1:] = cum[1:] - cum[:-1]                                     # :161-162
    q = widths / widths.sum()                                           # :195
    with np.errstate(divide="ignore", invalid="ignore"):
        logq = np.log(q)                                                # :196
        res = q * (logq - log_probs[: len(q)]) / 0.69315                # utils.py:33
    res[q == 0] = 0
    kl = float(res.sum())
    with np.errstate(invalid="ignore"):
        ent = probs_temp * log_probs_temp / 0.69315                     # utils.py:38
    ent[probs_temp == 0] = 0
    return float(log_probs[selection]), kl, float(-ent.sum())


def encode_step(row: np.ndarray, lo: int, hi: int, message: Sequence[int], cursor: int,
                *, temp: float, precision: int, topk: int,
                want_stats: bool = False):
    """One encode step (arithmetic.py:124-203) on one unmasked fp32 row.

    Returns ``(StepTrace, new_cursor, stats|None, cum_abs)``.
    """
    masked = mask_row(row, ENC_MASK)
    s, order = sort_desc(masked)
    probs = softmax_f64(s, temp)
    cur_range = hi - lo                                                 # :140
    cum, k = integer_cdf(probs, cur_range, topk)
    cum_abs = cum + lo                                                  # :165
    bits = list(message[cursor: cursor + precision])                    # :168
    if cursor + precision > len(message):
        bits = bits + [0] * (cursor + precision - len(message))         # :170
    message_idx = bits2int(reversed(bits))                              # :171
    selection = int(np.nonzero(cum_abs > message_idx)[0][0])            # :172
    new_bottom = int(cum_abs[selection - 1]) if selection > 0 else lo   # :175
    new_top = int(cum_abs[selection])                                   # :176
    n, nlo, nhi, _, _ = interval_update(new_bottom, new_top, precision)
    stats = _stats(s, temp, probs, cum, selection) if want_stats else None
    trace = StepTrace(token=int(order[selection]), selection=selection, k=k, nbits=n,
                      lo=nlo, hi=nhi, new_bottom=new_bottom, new_top=new_top)
    return trace, cursor + n, stats, cum_abs


def encode_stream(rows: Callable[[int], np.ndarray], message: Sequence[int], *,
                  temp: float = 1.0, precision: int = 16, topk: int = 50000,
                  max_steps: Optional[int] = None, want_stats: bool = False,
                  keep_trace: bool = True) -> EncodeResult:
    """The encode loop without ``finish_sent`` (arithmetic.py:96-217).

    ``rows(t)`` returns the fp32 logits row the model would emit at step ``t``.
    """
    message = [int(b) for b in message]
    lo, hi = 0, 1 << precision                                          # :96-98
    cursor, t = 0, 0
    tokens: List[int] = []
    trace: List[StepTrace] = []
    tot_lp = tot_kl = tot_h = 0.0
    while cursor < len(message):                                        # :114
        if max_steps is not None and t >= max_steps:
            break
        st, cursor, stats, _ = encode_step(rows(t), lo, hi, message, cursor, temp=temp,
                                           precision=precision, topk=topk, want_stats=want_stats)
        lo, hi = st.lo, st.hi
        tokens.append(st.token)
        if keep_trace:
            trace.append(st)
        if stats is not None:
            tot_lp += stats[0]; tot_kl += stats[1]; tot_h += stats[2]
        t += 1
    res = EncodeResult(tokens=tokens, bits_consumed=cursor, trace=trace)
    if want_stats and t > 0 and cursor > 0:
        res.avg_nll = -tot_lp / t                                       # :212
        res.avg_kl = tot_kl / t                                         # :213
        res.avg_hq = tot_h / t                                          # :214
        res.words_per_bit = t / cursor                                  # :215
    return res


def decode_step(row: np.ndarray, lo: int, hi: int, token: int, is_last: bool, *,
                temp: float, precision: int, topk: int):
    """One decode step (arithmetic.py:265-366).  Returns ``(StepTrace, bits, in_range)``.

    ``in_range`` is False when the observed token's rank is >= k; the reference
    then enters its tokenizer-specific BPE repair (:300-342), which is host
    string work outside the coder step -- the oracle falls back to rank 0 as
    the reference's last resort does (:342).
    """
    masked = mask_row(row, DEC_MASK)
    s, order = sort_desc(masked)
    probs = softmax_f64(s, temp)
    cur_range = hi - lo
    cum, k = integer_cdf(probs, cur_range, topk)
    cum_abs = cum + lo                                                  # :296
    rank = int(np.nonzero(order == token)[0][0])                        # :298
    in_range = rank < k
    if not in_range:
        rank = 0                                                        # :342
    new_bottom = int(cum_abs[rank - 1]) if rank > 0 else lo             # :347
    new_top = int(cum_abs[rank])                                        # :348
    n, nlo, nhi, bottom_bits, top_bits = interval_update(new_bottom, new_top, precision)
    bits = bottom_bits if is_last else top_bits[:n]                     # :356-359
    trace = StepTrace(token=int(token), selection=rank, k=k, nbits=n, lo=nlo, hi=nhi,
                      new_bottom=new_bottom, new_top=new_top)
    return trace, bits, in_range


def decode_stream(rows: Callable[[int], np.ndarray], tokens: Sequence[int], *,
                  temp: float = 1.0, precision: int = 16, topk: int = 50000,
                  keep_trace: bool = False):
    """The decode loop (arithmetic.py:246-373).  Returns ``(bits, trace)``."""
    lo, hi = 0, 1 << precision
    out: List[int] = []
    trace: List[StepTrace] = []
    for t, tok in enumerate(tokens):                                    # :255
        st, bits, _ = decode_step(rows(t), lo, hi, int(tok), t == len(tokens) - 1,
                                  temp=temp, precision=precision, topk=topk)
        lo, hi = st.lo, st.hi
        out += bits                                                     # :360
        if keep_trace:
            trace.append(st)
    return out, trace
