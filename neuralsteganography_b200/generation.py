"""Device-resident generation loop: GPT-2 trunk step + coder step, replayed as one CUDA graph.

Replaces the per-token Python loops of the reference -- code_base/arithmetic.py:114-210 (encode) / :255-371
(decode) for the arithmetic coder (A), src/neuralstego/codec/arithmetic.py:146-163 / :203-217 for the rank codec
(B) -- which synchronise with the host about ten times per token.  Here message bits, intervals, cursors, token
buffers and the KV cache live in HBM; the host replays a captured graph and reads one "all done" flag every
``poll_every`` steps.

Two decode schedules:
  * ``decode``            one trunk step + one coder step per cover token (the loop of the reference);
  * ``decode_prefill``    teacher-forced: the cover tokens are known, so the trunk runs over whole position tiles
                          (``[B, tile]`` tokens per call, tensor-core GEMMs with M = B * tile rows instead of B) and the
                          coder steps then walk the tile's logits.  The arithmetic decoder carries its interval from
                          step to step, the rank decoder carries nothing.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from .coder import ArithmeticStreams
from .codecs import CodecStreams
from .trunk import StaticGPT2


class StegoGenerator:
    """B streams sharing one context length, GPT-2 shaped trunk, coder (A) ``codec="ac"`` or (B) ``codec="rank"``."""

    def __init__(self, hf_model, batch: int, *, max_len: int = 256, precision: int = 16, temp: float = 1.0,
                 topk: int = 50000, finish_sent: bool = False, sent_end: Optional[torch.Tensor] = None,
                 device="cuda", use_graph: bool = True, trunk_dtype: torch.dtype = torch.float32,
                 collect_stats: bool = False, trunk_tf32: bool = False, codec: str = "ac",
                 codec_kw: Optional[dict] = None):
        if codec not in ("ac", "rank"):
            raise ValueError("codec must be 'ac' or 'rank'")
        self.B = int(batch)
        self.device = torch.device(device)
        self.trunk = StaticGPT2(hf_model, batch, max_len=max_len, device=device, dtype=trunk_dtype, tf32=trunk_tf32)
        self.V = self.trunk.vocab
        self.max_len = int(max_len)
        self.codec = codec
        if codec == "ac":
            self.kw = dict(precision=precision, temp=temp, topk=topk, finish_sent=finish_sent, sent_end=sent_end)
        else:
            self.kw = dict(temp=temp, **(codec_kw or {}))
        self.use_graph = bool(use_graph)
        self.logits = torch.zeros(self.B, self.V, dtype=torch.float32, device=self.device)
        self._rows = torch.arange(self.B, device=self.device)
        self.steps_run = 0
        self.collect_stats = bool(collect_stats)        # a6: NLL / KL / entropy sums (exact kernel, eager loop)
        self.stats_sum = torch.zeros(self.B, 3, dtype=torch.float64, device=self.device)
        self.stats_steps = 0
        self.history = None                             # rank codec: bits consumed per token [B, token_cap] (state["history"])
        self._coders = {}

    # one loop iteration: coder on the current logits, then the trunk on the token just fixed
    def _iter(self, coder, decode: bool, kv_len: Optional[int] = None) -> None:
        if decode:
            coder.decode_step(self.logits)
        else:
            if self.collect_stats:
                coding = (coder.phase == 0).to(torch.float64)[:, None]      # tail / finished streams add nothing
                coder.stats.zero_()
            coder.encode_step(self.logits)
            if self.collect_stats:
                self.stats_sum += coder.stats * coding
                self.stats_steps += int(coding.sum().item() > 0)
        last = (coder.ntok.long() - 1).clamp(min=0)
        if self.codec == "rank" and not decode:          # codec/arithmetic.py:159-166: per-token consumption
            self.history.scatter_(1, last[:, None], coder.nbits[:, None])
        prev = coder.tokens[self._rows, last].long().clamp(min=0)      # finished streams feed a stale token
        res = self.trunk.step(prev, kv_len, out=self.logits)           # fp32 trunk: lm_head writes self.logits, no [B, V] copy
        if res.data_ptr() != self.logits.data_ptr():
            self.logits.copy_(res)

    def _new_coder(self, token_cap: int, kw: dict):
        if self.codec == "ac":
            return ArithmeticStreams(self.B, self.V, device=self.device, token_cap=token_cap, **kw)
        return CodecStreams("rank", self.B, self.V, device=self.device, token_cap=token_cap, **kw)

    def _coder(self, decode: bool, token_cap: int, kw: dict):
        """Coder state (and with it the captured graphs) is kept across calls of the same shape."""
        key = (decode, int(token_cap))
        ent = self._coders.get(key)
        if ent is None:
            ent = {"coder": self._new_coder(token_cap, kw), "graphs": {}, "sig": None}
            self._coders = {k: v for k, v in self._coders.items() if k[0] != decode}   # one shape per direction
            self._coders[key] = ent
        return ent

    @staticmethod
    def _signature(c):
        return tuple(0 if t is None else t.data_ptr() for t in (c.msg, c.msg_len, c.tokens, c.ntok_total, c.out_bits, c.out_len,
                                                                 getattr(c, "total_bits", None)))

    def _run(self, ent: dict, decode: bool, max_steps: int, poll_every: int, ctx_len: int,
             on_poll: Optional[Callable[[object, int], bool]] = None) -> None:
        coder = ent["coder"]
        sig = self._signature(coder)
        if ent["sig"] != sig:                               # buffers moved: graphs captured on the old ones are stale
            ent["graphs"], ent["sig"] = {}, sig
        t = 0
        while t < max_steps:
            kv = self.trunk.kv_bucket(ctx_len + t + 1)      # host arithmetic only: the length is ctx_len + t
            graph = ent["graphs"].get(kv) if self.use_graph and not self.collect_stats else None
            if self.use_graph and not self.collect_stats and graph is None and (t >= 2 or ent["graphs"]):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._iter(coder, decode, kv)
                ent["graphs"][kv] = graph                   # capture does not execute: this step still has to run
            if graph is not None:
                graph.replay()
            else:
                self._iter(coder, decode, kv)
            t += 1
            if t % poll_every == 0:
                if coder.all_done():
                    break
                if on_poll is not None and on_poll(coder, t):
                    break
        self.steps_run = t

    def _prefill(self, contexts: torch.Tensor) -> None:
        contexts = contexts.to(self.device)
        if contexts.dim() == 1:
            contexts = contexts[None].expand(self.B, -1)
        contexts = contexts[:, -1022:]                                  # arithmetic.py:90
        self.trunk.reset()
        self.logits.copy_(self.trunk.prefill(contexts.contiguous()))

    def _room(self, contexts: torch.Tensor, max_tokens: Optional[int]) -> int:
        """Token budget of a stream: what fits the KV buffer, or ``max_tokens`` once the trunk slides its window."""
        room = self.max_len - min(int(contexts.shape[-1]), 1022) - 1
        if self.trunk.ring:
            return int(max_tokens) if max_tokens else max(room, 1)
        if max_tokens and max_tokens > room:
            raise ValueError("max_tokens=%d needs a KV buffer of at least 1023 slots (max_len >= 1023)" % max_tokens)
        return int(max_tokens) if max_tokens else room

    def _check_decode_room(self, ctx_len: int, n: int) -> None:
        """A plain (non-ring) KV buffer must hold the context and every cover token: the trunk would otherwise keep
        overwriting its last slot and produce wrong logits without any error."""
        if not self.trunk.ring and ctx_len + n + 1 > self.max_len:
            raise ValueError("decoding %d tokens after a %d-token context needs max_len >= %d (or >= 1023 for the "
                             "sliding window); this generator has max_len=%d" % (n, ctx_len, ctx_len + n + 1, self.max_len))

    def encode(self, contexts: torch.Tensor, messages: Sequence[Sequence[int]], *, poll_every: int = 16,
               max_tokens: Optional[int] = None, on_poll: Optional[Callable[[object, int], bool]] = None) -> List[List[int]]:
        """Cover tokens for one message (list of 0/1, in the codec's reading order) per stream."""
        room = self._room(contexts, max_tokens)
        ent = self._coder(False, max(1, room), self.kw)
        coder = ent["coder"]
        coder.set_messages(messages)
        if self.codec == "rank":
            if self.history is None or self.history.shape[1] != coder.token_cap:
                if self.history is not None:
                    ent["graphs"] = {}                      # the captured scatter wrote the old buffer
                self.history = torch.zeros(self.B, coder.token_cap, dtype=torch.uint8, device=self.device)
            self.history.zero_()
        if self.collect_stats:
            coder.stats = torch.zeros(self.B, 3, dtype=torch.float64, device=self.device)
            self.stats_sum.zero_(); self.stats_steps = 0
        self._prefill(contexts)
        self._run(ent, False, room, poll_every, min(int(contexts.shape[-1]), 1022), on_poll)
        self.coder = coder
        return coder.token_lists()

    def history_lists(self) -> List[List[int]]:
        """Rank codec: bits consumed per emitted token (``state["history"]``, codec/arithmetic.py:165-166)."""
        n = self.coder.ntok.cpu().numpy()
        h = self.history.cpu().numpy()
        return [h[r, : int(n[r])].astype(int).tolist() for r in range(self.B)]

    def _decode_coder(self, token_lists: Sequence[Sequence[int]], total_bits: Optional[Sequence[int]]):
        n = max((len(t) for t in token_lists), default=0)
        kw = dict(self.kw)
        if self.codec == "ac":
            kw["finish_sent"] = False
        ent = self._coder(True, max(1, n), kw)
        coder = ent["coder"]
        if self.codec == "ac":
            coder.set_tokens(token_lists)
        else:
            coder.set_tokens(token_lists, total_bits)
        return ent, coder, n

    def decode(self, contexts: torch.Tensor, token_lists: Sequence[Sequence[int]], *, poll_every: int = 16,
               total_bits: Optional[Sequence[int]] = None) -> List[List[int]]:
        """Recovered bits (message + trailing bits, as the reference returns them) per stream."""
        ctx_len = min(int(contexts.shape[-1]), 1022)
        ent, coder, n = self._decode_coder(token_lists, total_bits)
        self._check_decode_room(ctx_len, n)
        self._prefill(contexts)
        self._run(ent, True, n, poll_every, ctx_len)
        self.coder = coder
        return coder.bit_lists()

    def decode_prefill(self, contexts: torch.Tensor, token_lists: Sequence[Sequence[int]], *, tile: int = 32,
                       total_bits: Optional[Sequence[int]] = None) -> List[List[int]]:
        """Teacher-forced decode (BASELINE config 4): the trunk consumes the known cover in position tiles, the coder
        steps walk each tile's logits.  Same bits as :meth:`decode` whenever the tiled and the stepwise trunk produce
        the same logits (identical up to GEMM reduction order; pinned by the round trip of the caller)."""
        ctx_len = min(int(contexts.shape[-1]), 1022)
        ent, coder, n = self._decode_coder(token_lists, total_bits)
        if self.trunk.ring and ctx_len + n + 1 > self.trunk.ring:
            raise ValueError("decode_prefill serves covers that stay inside the 1022-token window; use decode()")
        self._check_decode_room(ctx_len, n)
        self._prefill(contexts)
        tok = coder.tokens.long().clamp(min=0)              # [B, cap]; padded slots feed token 0 (their logits are unused)
        done = 0
        coder.decode_step(self.logits)                      # token 0 is coded on the context's logits
        while done < n - 1:
            width = min(tile, n - 1 - done)
            # logits after tokens done .. done+width-1 (positions of cover tokens done+1 .. done+width)
            block = self.trunk.extend(tok[:, done:done + width])           # [B, width, V] fp32
            for j in range(width):
                coder.decode_step(block[:, j])
            done += width
        self.steps_run = n
        self.coder = coder
        return coder.bit_lists()
