"""Provider level on the GPU: framed chunks as one stream batch, batch pinning, the rank-codec provider with its
side-channel states, cover text -> spans with BPE repair, the research-script surface (statistics, finish_sent, <eos>)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from stub_tokenizer import StubTokenizer  # noqa: E402

from neuralsteganography_b200 import framing as F  # noqa: E402
from neuralsteganography_b200.exceptions import ConfigurationError  # noqa: E402

pytestmark = pytest.mark.gpu


def _model(vocab=2048, layers=2, width=64, heads=2, seed=1234):
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(seed)
    return GPT2LMHeadModel(GPT2Config(n_layer=layers, n_embd=width, n_head=heads, vocab_size=vocab)).eval().cuda()


def test_config1_with_ecc_all_chunks_in_one_batch_on_the_device():
    """stego_encode / stego_decode (api.py:707-807) with CRC + Reed-Solomon, every chunk a stream of one batch."""
    from neuralsteganography_b200.lm import B200ArithmeticLM, IdTokenizer
    lm = B200ArithmeticLM(_model(), IdTokenizer(2048), max_len=1024)
    message = bytes(np.random.default_rng(3).integers(0, 256, 300, dtype=np.uint8))
    q = {"temp": 0.9, "precision": 26, "topk": 300, "finish_sent": False}
    res = F.stego_encode(message, chunk_bytes=64, use_crc=True, ecc="rs", nsym=10, quality=q, seed_text="5 6 7", lm=lm, msg_id="m-1")
    assert len(res) == 5 and res.metadata.total == 5
    assert F.stego_decode(res, use_crc=True, ecc="rs", nsym=10, quality=q, seed_text="5 6 7", lm=lm) == message
    with pytest.raises(ConfigurationError):
        lm.encode_arithmetic([0] * 8, [1], quality={"top_p": 0.9})          # a rank-codec policy: rejected, not ignored


def test_decode_batch_is_pinned_to_the_encode_batch():
    """cuBLAS picks its algorithm by batch shape, so a cover encoded among 8 streams must be decoded with the
    same shapes: automatic inside one provider, ``batch_size=`` for a process that only decodes."""
    from neuralsteganography_b200.lm import B200ArithmeticLM, IdTokenizer, bytes_to_bits_lsb
    model = _model(vocab=50257, width=128, heads=4)
    q = {"temp": 0.9, "precision": 26, "topk": 300, "finish_sent": False}
    pk = [F.build_packet(bytes([r] * 20), msg_id="x", seq=r, total=8, cfg={"chunk_bytes": 20, "crc": True}) for r in range(8)]
    enc = B200ArithmeticLM(model, IdTokenizer(50257), max_len=1024)
    ctx = enc.encode_seed("1 2 3")
    covers = enc.encode_arithmetic_batch([bytes_to_bits_lsb(p) for p in pk], ctx, quality=q)
    one = enc.decode_arithmetic(covers[5], ctx, quality=q)                  # B=1 call, padded to 8 inside
    assert F.bits_to_bytes(one) == pk[5]
    dec = B200ArithmeticLM(model, IdTokenizer(50257), max_len=1024, batch_size=8)      # a fresh, decode-only provider
    assert F.bits_to_bytes(dec.decode_arithmetic(covers[2], ctx, quality=q)) == pk[2]
    with pytest.raises(ConfigurationError):
        dec.decode_arithmetic_batch(covers + covers, ctx, quality=q)       # more streams than the pinned batch


def test_rank_provider_roundtrip_states_and_oracle_history():
    """(B): what load_lm("gpt2-fa") runs in the reference -- ArithmeticLM over encode_with_lm/decode_with_lm -- on the
    device: tokens, {history, residual_bits} states (lm/arithmetic.py:186-190), FIFO decode, drain/load."""
    from neuralsteganography_b200.lm import B200RankLM, IdTokenizer, bytes_to_bits_lsb, bits_to_bytes_lsb
    from neuralsteganography_b200.trunk import StaticGPT2
    from oracle import codecs_oracle as K
    model = _model(vocab=42001, width=128, heads=4)
    lm = B200RankLM(model, IdTokenizer(42001), max_len=256)
    ctx = lm.encode_seed("9 8 7")
    payloads = [b"rank codec payload %d" % r for r in range(3)]
    q = {"temp": 1.0, "top_k": 4096}
    covers = lm.encode_arithmetic_batch([bytes_to_bits_lsb(p) for p in payloads], ctx, quality=q)
    states = lm.drain_states()
    assert len(states) == 3 and all(int.from_bytes(s["residual_bits"], "big") == 8 * len(p) for s, p in zip(states, payloads))
    assert all(sum(s["history"]) == 8 * len(p) and len(s["history"]) == len(c) for s, p, c in zip(states, payloads, covers))
    assert max(max(s["history"]) for s in states) == 12                    # floor(log2(4096)) bits per token
    back = lm.decode_arithmetic_batch(covers, ctx, quality=q)
    assert [bits_to_bytes_lsb(b) for b in back] == payloads
    with pytest.raises(ConfigurationError):
        lm.decode_arithmetic(covers[0], ctx, quality=q)                    # FIFO drained (lm/arithmetic.py:204-205)
    lm.load_states(states[1:2])
    assert bits_to_bytes_lsb(lm.decode_arithmetic(covers[1], ctx, quality=q)) == payloads[1]
    # identical logits -> the oracle's tokens and history (stream 0 replayed at the batch shape of the encode)
    tr = StaticGPT2(model, 3, max_len=256)
    c3 = torch.tensor(ctx)[None].expand(3, -1).cuda()
    rows = [tr.prefill(c3)[0].cpu().numpy()]
    for s in range(len(covers[0]) - 1):
        step = torch.tensor([c[min(s, len(c) - 1)] for c in covers], device="cuda")
        rows.append(tr.step(step, tr.kv_bucket(len(ctx) + s + 1))[0].cpu().numpy())
    want_tokens, want_history, want_total = K.rank_encode(lambda t: rows[t], payloads[0], temperature=1.0, top_k=4096)
    assert want_tokens == covers[0]
    assert tuple(want_history) == tuple(states[0]["history"]) and want_total == 8 * len(payloads[0])


def test_cover_text_to_spans_with_bpe_repair():
    """8f.2: seed + spans -> text -> re-tokenised (differently!) -> spans again, through the reference's repair heuristic."""
    from neuralsteganography_b200.lm import B200ArithmeticLM
    tok = StubTokenizer()
    lm = B200ArithmeticLM(_model(seed=7), tok, max_len=1024)
    q = {"temp": 1.0, "precision": 16, "topk": 6, "finish_sent": True}
    seed_text = "helloworld"
    message = b"meet at dawn"
    res = F.stego_encode(message, chunk_bytes=6, use_crc=True, ecc="none", quality=q, seed_text=seed_text, lm=lm, msg_id="t")
    spans = [list(s) for s in res]
    assert len(spans) == 2
    text = tok.decode(lm.encode_seed(seed_text) + [t for s in spans for t in s])
    retok = tok.encode(text)
    assert retok != tok.encode(seed_text) + [t for s in spans for t in s]   # the tokenizer does not round-trip
    got = lm.text_to_spans(text, seed_text, quality=q)
    assert got == spans
    assert F.stego_decode(got, use_crc=True, ecc="none", quality=q, seed_text=seed_text, lm=lm) == message


def test_code_base_surface_statistics_tail_and_text_roundtrip():
    from neuralsteganography_b200 import code_base as CB
    tok = StubTokenizer()
    model = _model(seed=11)
    ctx = tok.encode("thequickbrownfox")
    msg = np.random.default_rng(1).integers(0, 2, 96).tolist()
    for fn_e, fn_d, p in ((CB.encode_huffman, CB.decode_huffman, 3), (CB.encode_block, CB.decode_block, 3)):
        toks, nll, kl, wpb = fn_e(model, tok, msg, ctx, p, finish_sent=True)
        assert np.isfinite([nll, kl, wpb]).all() and nll > 0 and kl >= 0 and 0 < wpb <= 1
        last = tok.decode(toks[-1:])
        assert "." in last or "!" in last                                   # finish_sent tail (huffman_baseline.py:24,37-39)
        bits = fn_d(model, tok, toks, ctx, p)
        assert bits[: len(msg)] == msg
    # arithmetic: five-tuple with statistics, decode from the cover TEXT (re-tokenised, repaired)
    toks, nll, kl, wpb, hq = CB.encode_arithmetic(model, tok, msg, ctx, temp=0.9, precision=26, topk=6)
    assert np.isfinite([nll, kl, wpb, hq]).all()
    assert CB.decode_arithmetic(model, tok, tok.decode(toks), ctx, temp=0.9, precision=26, topk=6)[: len(msg)] == msg
    # run_single.py: text -> bits (AC decode at precision 40, topk 60000) -> cover -> bits -> text, <eos> stops the last leg
    out = CB.run_single(model, tok, "attackatdawn", context_tokens=ctx, mode="arithmetic", temp=0.9, precision=26, topk=6)
    assert out["reconstructed_text"].startswith("attackatdawn<eos>")
    assert out["recovered_bits"][: len(out["message_bits"])] == out["message_bits"]
    s_toks, s_nll, s_kl, s_hq = CB.sample(model, tok, 20, ctx, temperature=0.9, topk=50, seed=3)
    assert len(s_toks) == 20 and np.isfinite([s_nll, s_kl, s_hq]).all()


def test_cut_at_eos_is_the_reference_break():
    from neuralsteganography_b200.lm import cut_at_eos
    tok = StubTokenizer()
    ids = tok.encode("abc<eos>defgh")
    cut = cut_at_eos(ids, tok)
    assert tok.decode(cut).endswith("<eos>") and "<eos>" not in tok.decode(cut[:-1])
    assert cut_at_eos(tok.encode("abcdef"), tok) == tok.encode("abcdef")


def test_tensor_core_trunk_roundtrip_when_both_sides_set_it():
    """``trunk_tf32=True``: the trunk's GEMMs (lm_head above all) run on the tensor cores.  The logits differ from an fp32
    trunk's in their low bits, so the setting belongs to the channel: encoder and decoder both use it and the messages
    come back; the coder step itself is unchanged (same kernels, same integers for the logits it is given)."""
    from neuralsteganography_b200.lm import B200ArithmeticLM, IdTokenizer, bytes_to_bits_lsb
    model = _model(vocab=50257, width=128, heads=4)
    q = {"temp": 0.9, "precision": 26, "topk": 300, "finish_sent": False}
    pk = [F.build_packet(bytes([7 * r + 1] * 24), msg_id="t", seq=r, total=4, cfg={"chunk_bytes": 24, "crc": True}) for r in range(4)]
    lm = B200ArithmeticLM(model, IdTokenizer(50257), max_len=1024, trunk_tf32=True)
    ctx = lm.encode_seed("9 8 7")
    covers = lm.encode_arithmetic_batch([bytes_to_bits_lsb(p) for p in pk], ctx, quality=q)
    back = lm.decode_arithmetic_batch(covers, ctx, quality=q)
    assert [F.bits_to_bytes(b) for b in back] == pk
