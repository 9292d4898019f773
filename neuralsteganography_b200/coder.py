"""Device-resident batched arithmetic coder: B independent streams, one CTA each.

Host mirror of the loop in code_base/arithmetic.py:96-217 (encode) and
:246-373 (decode) of the reference, with the per-token work done by the CUDA
kernels in ``csrc/ns_coder.cu`` behind the C ABI of ``include/ns_coder.h``.
Message bits, cursors, intervals, token buffers and output bitstreams stay in
HBM; the host only enqueues steps and polls a done flag every few steps.
"""

from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N

NEWLINE2_ID = 628   # code_base/arithmetic.py:125


def pack_bits(bit_lists: Sequence[Sequence[int]], min_words: int = 1):
    """MSB-first packing into uint32 words -> (words [B, W] uint32, lengths [B] int32)."""
    B = len(bit_lists)
    lens = np.array([len(b) for b in bit_lists], dtype=np.int32)
    W = max(min_words, int((int(lens.max()) if B else 0) + 31) // 32 + 2)
    out = np.zeros((B, W), dtype=np.uint32)
    for r, bits in enumerate(bit_lists):
        if len(bits) == 0:
            continue
        arr = np.asarray(bits, dtype=np.uint8)
        by = np.packbits(arr)                       # first bit -> MSB of byte 0
        pad = (-len(by)) % 4
        if pad:
            by = np.concatenate([by, np.zeros(pad, dtype=np.uint8)])
        words = by.view(">u4").astype(np.uint32)
        out[r, : len(words)] = words
    return out, lens


def unpack_bits(words: np.ndarray, lens: np.ndarray) -> List[List[int]]:
    """Inverse of :func:`pack_bits`."""
    res: List[List[int]] = []
    for r in range(words.shape[0]):
        by = words[r].astype(">u4").view(np.uint8)
        bits = np.unpackbits(by)[: int(lens[r])]
        res.append(bits.astype(np.int64).tolist())
    return res


class ArithmeticStreams:
    """State of B arithmetic-coder streams on one GPU (SoA in HBM).

    Parameters follow ``encode_arithmetic`` of the reference
    (code_base/arithmetic.py:78-88): ``temp``, ``precision``, ``topk``,
    ``finish_sent``.  ``mask_ids`` defaults to the reference's forbidden tokens
    ``(V-1, 628)`` (:124-125).
    """

    def __init__(self, batch: int, vocab: int, *, precision: int = 16, temp: float = 1.0,
                 topk: int = 50000, finish_sent: bool = False, device="cuda",
                 mask_ids: Optional[Sequence[int]] = None, token_cap: int = 1024,
                 sent_end: Optional[torch.Tensor] = None, trace: bool = False, force_exact: bool = False,
                 variant: Optional[int] = None):
        self.lib = N.load()
        if not torch.cuda.is_available():
            raise N.NativeLibraryError("no CUDA device: the coder has no CPU fallback")
        if vocab > self.lib.ns_ac_max_vocab():
            raise N.NativeLibraryError("vocab %d exceeds kernel capacity %d" % (vocab, self.lib.ns_ac_max_vocab()))
        self.B, self.V = int(batch), int(vocab)
        self.precision, self.temp, self.topk = int(precision), float(temp), int(topk)
        self.finish_sent = bool(finish_sent)
        self.device = torch.device(device)
        if mask_ids is None:
            mask_ids = (vocab - 1, NEWLINE2_ID if vocab > NEWLINE2_ID else -1)
        self.mask_ids = tuple(int(x) for x in mask_ids)
        self.token_cap = int(token_cap)
        d = self.device
        self.lo = torch.zeros(self.B, dtype=torch.int64, device=d)
        self.hi = torch.full((self.B,), 1 << self.precision, dtype=torch.int64, device=d)
        self.phase = torch.zeros(self.B, dtype=torch.uint8, device=d)
        self.status = torch.zeros(self.B, dtype=torch.int32, device=d)
        self.ntok = torch.zeros(self.B, dtype=torch.int32, device=d)
        self.cursor = torch.zeros(self.B, dtype=torch.int32, device=d)
        self.nbits = torch.zeros(self.B, dtype=torch.uint8, device=d)
        self.tokens = torch.full((self.B, self.token_cap), -1, dtype=torch.int32, device=d)
        self.trace = torch.zeros((self.B, 4), dtype=torch.int64, device=d) if trace else None
        self.sent_end = sent_end
        self.force_exact = bool(force_exact)
        self.slow_ws = torch.zeros(self.B + 2, dtype=torch.int32, device=d)
        self.rank_ws = torch.zeros(self.B + 2, dtype=torch.int32, device=d)   # work list sweep kernel -> row-resident kernel
        # kernel choice of the throughput path: 0 = default (two rows in flight per SM for the threshold form of the
        # cutoff), 2 = always the single-row kernel (NS_AC_VARIANT overrides; results are identical)
        import os as _os
        self.variant = int(_os.environ.get("NS_AC_VARIANT", "0")) if variant is None else int(variant)
        self.msg = None
        self.msg_len = None
        self.ntok_total = None
        self.out_bits = None
        self.out_len = None
        self._params = N.AcParams()

    # ------------------------------------------------------------------ setup
    def reset(self) -> None:
        self.lo.zero_()
        self.hi.fill_(1 << self.precision)
        self.phase.zero_()
        self.status.zero_()
        self.ntok.zero_()
        self.cursor.zero_()

    def set_messages(self, bit_lists: Sequence[Sequence[int]]) -> None:
        """Load one message (list of 0/1) per stream for encoding."""
        assert len(bit_lists) == self.B
        words, lens = pack_bits(bit_lists)
        self.set_packed_messages(torch.from_numpy(words.view(np.int32)), torch.from_numpy(lens))

    def set_packed_messages(self, words_i32: torch.Tensor, lens_i32: torch.Tensor) -> None:
        # buffers are reused when the new messages fit, so that a captured CUDA graph stays valid across calls
        if self.msg is not None and self.msg.shape[0] == words_i32.shape[0] and self.msg.shape[1] >= words_i32.shape[1]:
            self.msg.zero_()
            self.msg[:, : words_i32.shape[1]].copy_(words_i32, non_blocking=True)
            self.msg_len.copy_(lens_i32, non_blocking=True)
        else:
            self.msg = words_i32.to(self.device, non_blocking=True).contiguous()
            self.msg_len = lens_i32.to(self.device, non_blocking=True).contiguous()
        self.reset()
        self.tokens.fill_(-1)
        # empty messages never enter the loop (arithmetic.py:114)
        self.phase.copy_(torch.where(self.msg_len > 0, 0, 2).to(torch.uint8))

    def set_tokens(self, token_lists: Sequence[Sequence[int]], out_bits_capacity: Optional[int] = None) -> None:
        """Load the observed cover tokens per stream for decoding."""
        assert len(token_lists) == self.B
        lens = np.array([len(t) for t in token_lists], dtype=np.int32)
        cap = max(1, int(lens.max()) if self.B else 1)
        if cap > self.token_cap:
            self.token_cap = cap
        tk = np.full((self.B, self.token_cap), -1, dtype=np.int32)
        for r, t in enumerate(token_lists):
            tk[r, : len(t)] = np.asarray(t, dtype=np.int32)
        self.set_token_tensor(torch.from_numpy(tk), torch.from_numpy(lens), out_bits_capacity)

    def set_token_tensor(self, tokens_i32: torch.Tensor, lens_i32: torch.Tensor,
                         out_bits_capacity: Optional[int] = None) -> None:
        # as in set_packed_messages: same shapes -> same buffers (CUDA graphs of the decode loop stay valid)
        if self.tokens is not None and tuple(self.tokens.shape) == tuple(tokens_i32.shape) and self.ntok_total is not None:
            self.tokens.copy_(tokens_i32, non_blocking=True)
            self.ntok_total.copy_(lens_i32, non_blocking=True)
        else:
            self.tokens = tokens_i32.to(self.device, non_blocking=True).contiguous()
            self.ntok_total = lens_i32.to(self.device, non_blocking=True).contiguous()
        self.token_cap = int(self.tokens.shape[1])
        self.reset()
        self.phase.copy_(torch.where(self.ntok_total > 0, 0, 2).to(torch.uint8))
        cap_bits = out_bits_capacity or (self.token_cap * self.precision + self.precision)
        words = (cap_bits + 31) // 32 + 2
        if self.out_bits is not None and tuple(self.out_bits.shape) == (self.B, words):
            self.out_bits.zero_(); self.out_len.zero_()
        else:
            self.out_bits = torch.zeros((self.B, words), dtype=torch.int32, device=self.device)
            self.out_len = torch.zeros(self.B, dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ steps
    def _fill_common(self, logits: torch.Tensor) -> N.AcParams:
        if logits.dtype != torch.float32 or logits.device.type != "cuda":
            raise N.NativeLibraryError("logits must be a float32 CUDA tensor")
        if logits.dim() != 2 or logits.shape[0] != self.B or logits.shape[1] != self.V or logits.stride(1) != 1:
            raise N.NativeLibraryError("logits must be [B, V] with unit inner stride")
        p = self._params
        p.logits = logits.data_ptr(); p.ld = logits.stride(0); p.B = self.B; p.V = self.V
        p.temp = self.temp; p.precision = self.precision; p.topk = self.topk
        p.mask_id[0] = self.mask_ids[0]; p.mask_id[1] = self.mask_ids[1] if len(self.mask_ids) > 1 else -1
        p.lo = self.lo.data_ptr(); p.hi = self.hi.data_ptr()
        p.phase = self.phase.data_ptr(); p.status = self.status.data_ptr()
        p.ntok = self.ntok.data_ptr(); p.token_cap = self.token_cap
        p.ntok_total = N.ptr(self.ntok_total)
        p.nbits_out = self.nbits.data_ptr()
        p.trace = N.ptr(self.trace)
        p.slow_ws = self.slow_ws.data_ptr()
        p.force_exact = int(self.force_exact)
        p.prof = N.ptr(getattr(self, "prof", None))
        p.stats = N.ptr(getattr(self, "stats", None))
        p.rank_ws = self.rank_ws.data_ptr(); p.scratch_stride = 0; p.scratch_slots = 0; p.variant = self.variant
        return p

    def encode_step(self, logits: torch.Tensor) -> None:
        """Enqueue one encode step on the current CUDA stream (no host sync)."""
        p = self._fill_common(logits)
        p.msg = self.msg.data_ptr(); p.msg_stride = self.msg.stride(0); p.msg_len = self.msg_len.data_ptr()
        p.cursor = self.cursor.data_ptr()
        p.token_out = self.tokens.data_ptr(); p.token_stride = self.tokens.stride(0)
        p.finish_sent = int(self.finish_sent); p.sent_end = N.ptr(self.sent_end)
        N.check(self.lib.ns_ac_encode_step(C.byref(p), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                "ns_ac_encode_step")

    def decode_step(self, logits: torch.Tensor) -> None:
        """Enqueue one decode step on the current CUDA stream (no host sync)."""
        p = self._fill_common(logits)
        p.token_in = self.tokens.data_ptr(); p.token_stride = self.tokens.stride(0)
        p.is_last = None
        p.out_bits = self.out_bits.data_ptr(); p.out_stride = self.out_bits.stride(0)
        p.out_len = self.out_len.data_ptr()
        N.check(self.lib.ns_ac_decode_step(C.byref(p), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                "ns_ac_decode_step")

    def debug_bins(self, logits: torch.Tensor):
        """Integer bin widths per token id and (k, slack, total, range) per stream (parity tests)."""
        p = self._fill_common(logits)
        q = torch.zeros((self.B, self.V), dtype=torch.int64, device=self.device)
        meta = torch.zeros((self.B, 4), dtype=torch.int64, device=self.device)
        N.check(self.lib.ns_ac_debug_bins(C.byref(p), C.c_void_p(q.data_ptr()), C.c_void_p(meta.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ns_ac_debug_bins")
        return q, meta

    # ------------------------------------------------------------------ loops
    def all_done(self) -> bool:
        """One tiny device->host read; call it every few steps, not every token."""
        return bool((self.phase == N.PHASE_DONE).all().item())

    def encode(self, logits_fn: Callable[[int], torch.Tensor], *, max_steps: Optional[int] = None,
               poll_every: int = 16) -> List[List[int]]:
        """Run the encode loop; ``logits_fn(t)`` returns the [B, V] logits of step t."""
        steps = max_steps if max_steps is not None else self.token_cap
        for t in range(steps):
            self.encode_step(logits_fn(t))
            if (t + 1) % poll_every == 0 and self.all_done():
                break
        return self.token_lists()

    def decode(self, logits_fn: Callable[[int], torch.Tensor], *, poll_every: int = 16) -> List[List[int]]:
        steps = int(self.ntok_total.max().item()) if self.B else 0
        for t in range(steps):
            self.decode_step(logits_fn(t))
        return self.bit_lists()

    # ------------------------------------------------------------------ results
    def token_lists(self) -> List[List[int]]:
        n = self.ntok.cpu().numpy()
        tk = self.tokens.cpu().numpy()
        return [tk[r, : int(n[r])].tolist() for r in range(self.B)]

    def bit_lists(self) -> List[List[int]]:
        words = self.out_bits.cpu().numpy().view(np.uint32)
        return unpack_bits(words, self.out_len.cpu().numpy())
