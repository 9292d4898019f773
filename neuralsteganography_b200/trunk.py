"""GPT-2 trunk with a static KV cache: the LM side of the device-resident generation loop.

The reference drives HuggingFace ``GPT2LMHeadModel`` one token at a time and converts its
cache object every step (code_base/arithmetic.py:115-122, utils.py:19-30).  Here the same
weights run through plain PyTorch ops over pre-allocated KV buffers so that one decoding
step has fixed shapes and no host decisions -- it can be captured in a CUDA graph together
with the coder step.  PyTorch is plumbing here (GEMMs via cuBLAS); the product is the coder.

Position rule of the reference: the first call sees the whole context at positions
0..L-1, later calls use ``past_len % n_positions`` (arithmetic.py:44-48); the context is
cut to its last 1022 tokens (:90).  After every call the reference keeps the last 1022 cache
entries (``limit_past``, utils.py:19-30): a token therefore attends to at most 1022 earlier
entries plus itself, and once the cache is full every new token sits at position
``1022 % n_positions``.  With buffers of at least 1023 slots the KV cache here is a ring of
1023 slots (attention does not depend on the order of the keys), so streams of any length
follow the same rule; shorter buffers only serve streams that end before they fill.
"""

from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


class _matmul_tf32:
    """Scoped ``torch.backends.cuda.matmul.allow_tf32`` (a host-side switch: also effective during graph capture)."""

    def __init__(self, on: bool):
        self.on = on

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        if self.on:
            torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


class StaticGPT2:
    """Inference-only GPT-2 (gelu_new, pre-LN) over static KV buffers ``[layers, B, heads, T, hd]``."""

    def __init__(self, hf_model, batch: int, max_len: Optional[int] = None, device="cuda",
                 dtype: torch.dtype = torch.float32, tf32: bool = False):
        cfg = hf_model.config
        self.n_layer, self.n_head, self.n_embd = cfg.n_layer, cfg.n_head, cfg.n_embd
        self.n_positions, self.vocab = cfg.n_positions, cfg.vocab_size
        self.eps = cfg.layer_norm_epsilon
        self.hd = self.n_embd // self.n_head
        self.B = int(batch)
        self.T = int(max_len or cfg.n_positions)
        self.device, self.dtype = torch.device(device), dtype
        # fp32 weights with TF32 tensor-core GEMMs (PyTorch's default is strict fp32, as in the reference): ~4x
        # faster linear layers at ~1e-3 relative error in the logits; encoder and decoder must use the same setting
        self.tf32 = bool(tf32)
        sd = {k: v.detach().to(self.device, dtype) for k, v in hf_model.state_dict().items()}
        g = lambda name: sd[name].contiguous()
        self.wte, self.wpe = g("transformer.wte.weight"), g("transformer.wpe.weight")
        self.layers = []
        for i in range(self.n_layer):
            p = "transformer.h.%d." % i
            self.layers.append(dict(
                ln1w=g(p + "ln_1.weight"), ln1b=g(p + "ln_1.bias"),
                qkvw=g(p + "attn.c_attn.weight"), qkvb=g(p + "attn.c_attn.bias"),       # Conv1D: x @ W + b
                pw=g(p + "attn.c_proj.weight"), pb=g(p + "attn.c_proj.bias"),
                ln2w=g(p + "ln_2.weight"), ln2b=g(p + "ln_2.bias"),
                fcw=g(p + "mlp.c_fc.weight"), fcb=g(p + "mlp.c_fc.bias"),
                ow=g(p + "mlp.c_proj.weight"), ob=g(p + "mlp.c_proj.bias")))
        self.lnfw, self.lnfb = g("transformer.ln_f.weight"), g("transformer.ln_f.bias")
        self.lm_head = g("lm_head.weight") if "lm_head.weight" in sd else self.wte   # tied
        self.k = torch.zeros(self.n_layer, self.B, self.n_head, self.T, self.hd, device=self.device, dtype=dtype)
        self.v = torch.zeros_like(self.k)
        self.length = torch.zeros((), dtype=torch.long, device=self.device)   # tokens in the cache (same for all streams)
        self._arange_t = torch.arange(self.T, device=self.device)
        self._host_len = 0                                  # host copy of `length` while only eager calls touched it
        self.window = 1022                                  # utils.py:19-30 (hard-coded in the reference)
        self.ring = self.window + 1 if self.T >= self.window + 1 else 0   # ring slots; 0 = plain buffer

    # ------------------------------------------------------------------ context (variable length, eager)
    @torch.no_grad()
    def prefill(self, context: torch.Tensor) -> torch.Tensor:
        """Run ``context`` [B, L] (L <= T - 1) through the trunk; returns fp32 logits [B, V] of its last token."""
        with _matmul_tf32(self.tf32):
            return self._prefill(context)

    def _prefill(self, context: torch.Tensor) -> torch.Tensor:
        B, L = context.shape
        assert B == self.B and L < self.T
        pos = torch.arange(L, device=self.device)
        x = self.wte[context] + self.wpe[pos][None]
        mask = torch.ones(L, L, device=self.device, dtype=torch.bool).tril()
        for i, w in enumerate(self.layers):
            h = F.layer_norm(x, (self.n_embd,), w["ln1w"], w["ln1b"], self.eps)
            qkv = h @ w["qkvw"] + w["qkvb"]
            q, k, v = qkv.split(self.n_embd, dim=-1)
            q = q.view(B, L, self.n_head, self.hd).transpose(1, 2)
            k = k.view(B, L, self.n_head, self.hd).transpose(1, 2)
            v = v.view(B, L, self.n_head, self.hd).transpose(1, 2)
            self.k[i, :, :, :L] = k
            self.v[i, :, :, :L] = v
            att = (q @ k.transpose(-1, -2)) / math.sqrt(self.hd)
            att = att.masked_fill(~mask, torch.finfo(att.dtype).min).softmax(-1)
            a = (att @ v).transpose(1, 2).reshape(B, L, self.n_embd)
            x = x + (a @ w["pw"] + w["pb"])
            h = F.layer_norm(x, (self.n_embd,), w["ln2w"], w["ln2b"], self.eps)
            x = x + (F.gelu(h @ w["fcw"] + w["fcb"], approximate="tanh") @ w["ow"] + w["ob"])
        self.length.fill_(L)
        self._host_len = L
        x = F.layer_norm(x[:, -1], (self.n_embd,), self.lnfw, self.lnfb, self.eps)
        return (x @ self.lm_head.t()).float().contiguous()

    # ------------------------------------------------------------------ a tile of known tokens (teacher forcing, eager)
    @torch.no_grad()
    def extend(self, tokens: torch.Tensor) -> torch.Tensor:
        """Append the known tokens ``[B, W]`` at positions ``length .. length+W-1`` in one pass (causal inside the
        tile, full attention to the cache); returns fp32 logits ``[B, W, V]`` -- entry ``j`` is the distribution of the
        token that follows ``tokens[:, j]``.  Serves covers that stay inside the KV buffer / the 1022-token window."""
        with _matmul_tf32(self.tf32):
            return self._extend(tokens)

    def _extend(self, tokens: torch.Tensor) -> torch.Tensor:
        B, W = tokens.shape
        L = self._host_len if self._host_len is not None else int(self.length.item())
        limit = self.ring if self.ring else self.T
        if B != self.B or L + W > limit:
            raise ValueError("extend: %d cached + %d new tokens exceed the %d usable KV slots" % (L, W, limit))
        pos = torch.arange(L, L + W, device=self.device).remainder(self.n_positions)     # arithmetic.py:44-48
        x = self.wte[tokens] + self.wpe[pos][None]
        # Linear layers run on all B*W rows at once.  Attention runs per position with exactly the shapes of `_step`
        # (one query against the live KV bucket): cuBLAS' batched GEMM is not row-wise identical between 1 and W queries
        # (tests/diag/gemm_invariance.py), and encoder and decoder must see bit-identical logits.
        for i, w in enumerate(self.layers):
            h = F.layer_norm(x, (self.n_embd,), w["ln1w"], w["ln1b"], self.eps)
            qkv = h @ w["qkvw"] + w["qkvb"]
            q, k, v = qkv.split(self.n_embd, dim=-1)
            self.k[i][:, :, L:L + W] = k.view(B, W, self.n_head, self.hd).transpose(1, 2)
            self.v[i][:, :, L:L + W] = v.view(B, W, self.n_head, self.hd).transpose(1, 2)
            outs = []
            for j in range(W):
                Tk = self.kv_bucket(L + j + 1)
                qj = q[:, j].reshape(B, self.n_head, 1, self.hd)
                kk, vv = self.k[i][:, :, :Tk], self.v[i][:, :, :Tk]
                live = (self._arange_t[:Tk] <= L + j)[None, None, None, :]
                att = (qj @ kk.transpose(-1, -2)) / math.sqrt(self.hd)
                att = att.masked_fill(~live, torch.finfo(att.dtype).min).softmax(-1)
                outs.append((att @ vv).reshape(B, self.n_embd))
            a = torch.stack(outs, dim=1)
            x = x + (a @ w["pw"] + w["pb"])
            h = F.layer_norm(x, (self.n_embd,), w["ln2w"], w["ln2b"], self.eps)
            x = x + (F.gelu(h @ w["fcw"] + w["fcb"], approximate="tanh") @ w["ow"] + w["ob"])
        self.length.fill_(L + W)
        self._host_len = L + W
        x = F.layer_norm(x, (self.n_embd,), self.lnfw, self.lnfb, self.eps)
        return (x @ self.lm_head.t()).float()

    # ------------------------------------------------------------------ one token (fixed shapes, graph-capturable)
    @torch.no_grad()
    def kv_bucket(self, needed: int) -> int:
        """Smallest power-of-two prefix of the KV buffer (>= 64 slots) that holds ``needed`` entries; the whole
        buffer once the ring is in use.  A decoding step only reads that prefix, so a graph per bucket keeps the
        attention traffic proportional to the live length instead of the buffer size."""
        if self.ring and needed > self.ring:
            return self.T
        b = 64
        while b < needed:
            b *= 2
        return min(b, self.T)

    def step(self, tokens: torch.Tensor, kv_len: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Append ``tokens`` [B] (int64) at position ``length``; returns fp32 logits [B, V].  No host sync.
        ``kv_len`` (host int, >= length + 1): only that prefix of the KV buffers is attended to."""
        with _matmul_tf32(self.tf32):
            return self._step(tokens, kv_len, out)

    def _step(self, tokens: torch.Tensor, kv_len: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = self.B
        Tk = self.T if kv_len is None else int(kv_len)
        ar = self._arange_t[:Tk]
        if self.ring:
            pos = self.length.clamp(max=self.window).remainder(self.n_positions).view(1)   # cache length, then :44-48
            slot = self.length.remainder(self.ring).view(1)                          # overwrites the entry limit_past dropped
            live = ((ar <= self.length) & (ar < self.ring))[None, None, None, :]
        else:
            pos = self.length.remainder(self.n_positions).view(1)                    # arithmetic.py:44-48
            slot = self.length.clamp(max=self.T - 1).view(1)                         # device index: no host sync
            live = (ar <= self.length)[None, None, None, :]                          # keys 0..length
        x = self.wte.index_select(0, tokens) + self.wpe.index_select(0, pos)         # tensor indices: no host sync
        for i, w in enumerate(self.layers):
            h = F.layer_norm(x, (self.n_embd,), w["ln1w"], w["ln1b"], self.eps)
            qkv = h @ w["qkvw"] + w["qkvb"]
            q, k, v = qkv.split(self.n_embd, dim=-1)
            q = q.view(B, self.n_head, 1, self.hd)
            k = k.view(B, self.n_head, 1, self.hd)
            v = v.view(B, self.n_head, 1, self.hd)
            self.k[i].index_copy_(2, slot, k)                                        # cache row `length`
            self.v[i].index_copy_(2, slot, v)
            kk, vv = self.k[i][:, :, :Tk], self.v[i][:, :, :Tk]                      # views: the live prefix
            att = (q @ kk.transpose(-1, -2)) / math.sqrt(self.hd)                    # [B, H, 1, Tk]
            att = att.masked_fill(~live, torch.finfo(att.dtype).min).softmax(-1)
            a = (att @ vv).reshape(B, self.n_embd)
            x = x + (a @ w["pw"] + w["pb"])
            h = F.layer_norm(x, (self.n_embd,), w["ln2w"], w["ln2b"], self.eps)
            x = x + (F.gelu(h @ w["fcw"] + w["fcb"], approximate="tanh") @ w["ow"] + w["ob"])
        self.length.add_(1)
        self._host_len = None                               # replayed under CUDA graphs: the host no longer knows the length
        x = F.layer_norm(x, (self.n_embd,), self.lnfw, self.lnfb, self.eps)
        if out is not None and self.dtype == torch.float32:       # the lm_head GEMM writes the coder's logits buffer directly
            return torch.matmul(x, self.lm_head.t(), out=out)
        return (x @ self.lm_head.t()).float().contiguous()

    def reset(self) -> None:
        self.length.zero_()
        self._host_len = 0
