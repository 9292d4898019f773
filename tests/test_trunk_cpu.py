"""The trunk's sliding KV window against a naive restatement of the reference's cache rule
(code_base/utils.py:19-30 ``limit_past`` + code_base/arithmetic.py:44-48 position rule).  Pure PyTorch, CPU."""
import math

import torch
import torch.nn.functional as F

from neuralsteganography_b200.trunk import StaticGPT2


def _tiny(n_positions=1024):
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(7)
    cfg = GPT2Config(vocab_size=96, n_positions=n_positions, n_embd=16, n_layer=2, n_head=2)
    return GPT2LMHeadModel(cfg).eval()


class NaiveWindowed:
    """Chronological KV lists, cut to the last 1022 entries after every call, position = cache length % n_positions."""

    def __init__(self, trunk: StaticGPT2):
        self.t = trunk
        self.k = [None] * trunk.n_layer
        self.v = [None] * trunk.n_layer

    def call(self, ids: torch.Tensor) -> torch.Tensor:          # ids [B, n]
        t = self.t
        B, n = ids.shape
        past = 0 if self.k[0] is None else self.k[0].shape[2]
        pos = (torch.arange(n) if past == 0 else torch.tensor([past % t.n_positions]))
        x = t.wte[ids] + t.wpe[pos][None]
        for i, w in enumerate(t.layers):
            h = F.layer_norm(x, (t.n_embd,), w["ln1w"], w["ln1b"], t.eps)
            q, k, v = (h @ w["qkvw"] + w["qkvb"]).split(t.n_embd, dim=-1)
            sh = lambda z: z.view(B, n, t.n_head, t.hd).transpose(1, 2)
            q, k, v = sh(q), sh(k), sh(v)
            if self.k[i] is not None:
                k = torch.cat([self.k[i], k], dim=2)
                v = torch.cat([self.v[i], v], dim=2)
            att = (q @ k.transpose(-1, -2)) / math.sqrt(t.hd)
            L = k.shape[2]
            mask = torch.ones(n, L, dtype=torch.bool).tril(L - n)
            att = att.masked_fill(~mask, torch.finfo(att.dtype).min).softmax(-1)
            a = (att @ v).transpose(1, 2).reshape(B, n, t.n_embd)
            x = x + (a @ w["pw"] + w["pb"])
            h = F.layer_norm(x, (t.n_embd,), w["ln2w"], w["ln2b"], t.eps)
            x = x + (F.gelu(h @ w["fcw"] + w["fcb"], approximate="tanh") @ w["ow"] + w["ob"])
            self.k[i], self.v[i] = k[:, :, -1022:], v[:, :, -1022:]          # limit_past
        x = F.layer_norm(x[:, -1], (t.n_embd,), t.lnfw, t.lnfb, t.eps)
        return x @ t.lm_head.t()


def test_ring_cache_follows_the_1022_window():
    model = _tiny()
    B, L, steps = 2, 1000, 60                                   # the cache fills at step 22, then slides
    trunk = StaticGPT2(model, B, max_len=1024, device="cpu")
    assert trunk.ring == 1023
    naive = NaiveWindowed(trunk)
    g = torch.Generator().manual_seed(3)
    ctx = torch.randint(0, 96, (B, L), generator=g)
    a, b = trunk.prefill(ctx), naive.call(ctx)
    assert torch.allclose(a, b, atol=2e-5)
    for s in range(steps):
        tok = torch.randint(0, 96, (B,), generator=g)
        a, b = trunk.step(tok), naive.call(tok[:, None])
        assert torch.allclose(a, b, atol=2e-5), s
    assert int(trunk.length) == L + steps


def test_short_buffer_matches_until_it_fills():
    model = _tiny()
    B, L, steps = 2, 5, 20
    trunk = StaticGPT2(model, B, max_len=32, device="cpu")
    assert trunk.ring == 0
    naive = NaiveWindowed(trunk)
    g = torch.Generator().manual_seed(4)
    ctx = torch.randint(0, 96, (B, L), generator=g)
    assert torch.allclose(trunk.prefill(ctx), naive.call(ctx), atol=2e-5)
    for s in range(steps):
        tok = torch.randint(0, 96, (B,), generator=g)
        assert torch.allclose(trunk.step(tok), naive.call(tok[:, None]), atol=2e-5), s


def test_kv_bucket_prefix_gives_the_same_logits():
    """Attending to the live power-of-two prefix of the KV buffer (what the captured graphs do) changes nothing."""
    model = _tiny()
    B, L = 2, 40
    a = StaticGPT2(model, B, max_len=256, device="cpu")
    b = StaticGPT2(model, B, max_len=256, device="cpu")
    assert [a.kv_bucket(n) for n in (1, 64, 65, 128, 129, 300)] == [64, 64, 128, 128, 256, 256]
    g = torch.Generator().manual_seed(5)
    ctx = torch.randint(0, 96, (B, L), generator=g)
    assert torch.allclose(a.prefill(ctx), b.prefill(ctx))
    for s in range(40):                                           # crosses the 64-slot bucket at step 24
        tok = torch.randint(0, 96, (B,), generator=g)
        assert torch.allclose(a.step(tok), b.step(tok, b.kv_bucket(L + s + 1)), atol=1e-6), s
    ring = StaticGPT2(model, B, max_len=1024, device="cpu")
    assert ring.kv_bucket(1023) == 1024 and ring.kv_bucket(1024) == 1024 and ring.kv_bucket(5000) == 1024


def test_extend_tiles_match_token_steps():
    """Teacher-forced tiles (``extend``) give the logits of the one-token steps (config 4's batched decode)."""
    model = _tiny()
    B, L, n = 2, 5, 11
    g = torch.Generator().manual_seed(9)
    ctx = torch.randint(0, 96, (B, L), generator=g)
    toks = torch.randint(0, 96, (B, n), generator=g)
    a = StaticGPT2(model, B, max_len=64, device="cpu")
    b = StaticGPT2(model, B, max_len=64, device="cpu")
    a.prefill(ctx); b.prefill(ctx)
    steps = torch.stack([a.step(toks[:, j]) for j in range(n)], dim=1)          # [B, n, V]
    tiles = torch.cat([b.extend(toks[:, :4]), b.extend(toks[:, 4:5]), b.extend(toks[:, 5:])], dim=1)
    assert tiles.shape == steps.shape
    assert torch.allclose(tiles, steps, atol=1e-5, rtol=1e-5)
    assert int(b.length) == L + n
    import pytest
    with pytest.raises(ValueError):
        b.extend(torch.zeros(B, 64, dtype=torch.long))
