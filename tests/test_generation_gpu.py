"""Device-resident generation loop (trunk + coder under a CUDA graph) and the provider boundary on the GPU."""
import json

import numpy as np
import pytest
import torch

from oracle import ac_oracle as O
from oracle.inputs import message_bits

pytestmark = pytest.mark.gpu


def _small_model(vocab=50257, layers=2, width=128, heads=4):
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(1234)
    return GPT2LMHeadModel(GPT2Config(n_layer=layers, n_embd=width, n_head=heads, vocab_size=vocab)).eval().cuda()


def _replay_logits(model, ctx, tokens, max_len=128):
    """fp32 logits the trunk produced at each step of one stream (teacher forced, eager)."""
    from neuralsteganography_b200.trunk import StaticGPT2
    tr = StaticGPT2(model, 1, max_len=max_len)
    rows = [tr.prefill(ctx[None].cuda())[0].cpu().numpy()]
    for s, t in enumerate(tokens[:-1]):                      # same KV prefix per step as the generation loop
        rows.append(tr.step(torch.tensor([t], device="cuda"), tr.kv_bucket(len(ctx) + s + 1))[0].cpu().numpy())
    return rows


def test_static_trunk_matches_huggingface():
    from neuralsteganography_b200.trunk import StaticGPT2
    model = _small_model(vocab=2048)
    B = 3
    tr = StaticGPT2(model, B, max_len=32)
    ctx = torch.randint(0, 2048, (B, 5), device="cuda")
    with torch.no_grad():
        out = model(ctx, use_cache=True)
        assert torch.allclose(out.logits[:, -1].float(), tr.prefill(ctx), atol=2e-4, rtol=1e-4)
        past = out.past_key_values
        for _ in range(4):
            tok = torch.randint(0, 2048, (B,), device="cuda")
            out = model(tok[:, None], past_key_values=past, use_cache=True)
            past = out.past_key_values
            assert torch.allclose(out.logits[:, -1].float(), tr.step(tok), atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("use_graph", [False, True])
def test_generation_loop_roundtrip_and_oracle_parity(use_graph):
    """config 2 in miniature: GPT-2 shaped random-init trunk, temp 0.9, precision 26, topk 300."""
    from neuralsteganography_b200.generation import StegoGenerator
    model = _small_model()
    B = 4
    gen = StegoGenerator(model, B, max_len=128, precision=26, temp=0.9, topk=300, use_graph=use_graph)
    ctx = torch.tensor([50256, 11, 22])
    msgs = [message_bits(10 + r, 160 - 8 * r).tolist() for r in range(B)]
    toks = gen.encode(ctx, msgs, poll_every=4)
    assert gen.coder.all_done()
    bits = gen.decode(ctx, toks, poll_every=4)
    for r in range(B):
        assert bits[r][: len(msgs[r])] == msgs[r], r
    # identical logits -> identical tokens: replay stream 0 through the CPU oracle
    gen1 = StegoGenerator(model, 1, max_len=128, precision=26, temp=0.9, topk=300, use_graph=use_graph)
    t1 = gen1.encode(ctx, [msgs[0]], poll_every=4)[0]
    rows = _replay_logits(model, ctx, t1)
    ref = O.encode_stream(lambda t: rows[t], msgs[0], temp=0.9, precision=26, topk=300, max_steps=len(t1))
    assert ref.tokens == t1
    back, _ = O.decode_stream(lambda t: rows[t], t1, temp=0.9, precision=26, topk=300)
    assert gen1.decode(ctx, [t1])[0] == back


def test_graph_and_eager_loops_agree():
    from neuralsteganography_b200.generation import StegoGenerator
    model = _small_model()
    ctx = torch.tensor([50256, 5, 6, 7])
    msgs = [message_bits(77 + r, 120).tolist() for r in range(3)]
    outs = []
    for use_graph in (False, True):
        gen = StegoGenerator(model, 3, max_len=96, precision=16, temp=1.0, topk=50000, use_graph=use_graph)
        outs.append(gen.encode(ctx, msgs, poll_every=2))
    assert outs[0] == outs[1]


def test_provider_roundtrips_a_framed_packet():
    from neuralsteganography_b200.lm import B200ArithmeticLM, IdTokenizer, bits_to_bytes_lsb, bytes_to_bits_lsb
    model = _small_model()
    lm = B200ArithmeticLM(model, IdTokenizer(50257), max_len=512)
    pkt = json.dumps({"cfg": {"chunk_bytes": 32, "crc": True, "ecc": "none", "nsym": 0}, "msg_id": "ab12", "payload": "aGVsbG8gd29ybGQ=",
                      "seq": 0, "total": 1, "version": 1}, separators=(",", ":"), sort_keys=True).encode()
    ctx = lm.encode_seed("7 8 9")
    assert ctx[0] == 50256
    quality = {"temp": 0.9, "precision": 26, "topk": 300, "finish_sent": False}
    toks = lm.encode_arithmetic(bytes_to_bits_lsb(pkt), ctx, quality=quality)
    bits = lm.decode_arithmetic(toks, ctx, quality=quality)
    assert bits_to_bytes_lsb(bits) == pkt
    many = lm.encode_arithmetic_batch([bytes_to_bits_lsb(pkt)] * 3, ctx, quality=quality)
    assert many[0] == many[1] == many[2]
    back = lm.decode_arithmetic_batch(many, ctx, quality=quality)
    assert all(bits_to_bytes_lsb(b) == pkt for b in back)


def test_generation_slides_the_1022_token_window():
    """a7: a stream that outgrows the KV window (context 1012 + 48 tokens) keeps coding under the CUDA graph:
    ring cache + position rule of code_base/arithmetic.py:44-48 / utils.py:19-30; parity with the oracle on the
    logits the windowed trunk produced, and the message comes back."""
    from neuralsteganography_b200.generation import StegoGenerator
    from neuralsteganography_b200.trunk import StaticGPT2
    model = _small_model(vocab=2048, layers=2, width=64, heads=2)
    B = 2
    gen = StegoGenerator(model, B, max_len=1024, precision=16, temp=1.0, topk=2048, use_graph=True)
    assert gen.trunk.ring == 1023
    g = torch.Generator().manual_seed(11)
    ctx = torch.randint(0, 2047, (1012,), generator=g)
    msgs = [message_bits(70 + r, 400).tolist() for r in range(B)]
    toks = gen.encode(ctx, msgs, poll_every=8, max_tokens=200)
    assert gen.coder.all_done()
    assert min(len(t) for t in toks) > 1023 - 1012                     # the window did slide
    bits = gen.decode(ctx, toks, poll_every=8)
    for r in range(B):
        assert bits[r][: len(msgs[r])] == msgs[r], r
    tr = StaticGPT2(model, 1, max_len=1024)
    rows = [tr.prefill(ctx[None].cuda())[0].cpu().numpy()]
    for s, t in enumerate(toks[0][:-1]):
        rows.append(tr.step(torch.tensor([t], device="cuda"), tr.kv_bucket(len(ctx) + s + 1))[0].cpu().numpy())
    res = O.encode_stream(lambda t: rows[t], msgs[0], temp=1.0, precision=16, topk=2048, max_steps=len(rows))
    assert res.tokens == toks[0]


def test_config4_shape_batched_decode_recovers_every_message():
    """BASELINE config 4: gpt2-fa-shaped random-init trunk (42001 tokens, 12 layers x 768), 1024 streams, cover
    produced by our own encode, then the batched decode path recovers every message from the cover token ids."""
    from neuralsteganography_b200.generation import StegoGenerator
    from neuralsteganography_b200.lm import random_init_model
    _tok, model = random_init_model("gpt2-fa")
    model = model.cuda()
    B, bits = 1024, 768
    gen = StegoGenerator(model, B, max_len=320, precision=26, temp=0.9, topk=300, use_graph=True)
    assert gen.V == 42001
    ctx = torch.tensor([5, 11, 22])
    rng = np.random.default_rng(44)
    msgs = [rng.integers(0, 2, bits - 8 * (r % 5)).tolist() for r in range(B)]
    toks = gen.encode(ctx, msgs, poll_every=16)
    assert gen.coder.all_done()
    assert int((gen.coder.status & 0xB).sum().item()) == 0
    assert max(len(t) for t in toks) <= 256
    got = gen.decode(ctx, toks, poll_every=16)
    for r in range(B):
        assert got[r][: len(msgs[r])] == msgs[r], r
