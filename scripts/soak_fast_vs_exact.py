"""Soak: the throughput kernel against the exact kernel over many rows, ranges and settings (both forms of the
cutoff, encode and decode).  Prints the number of row-steps compared and of differences (expected: 0; the
floating-point contract of DESIGN.md section 4 puts the rate near 1e-7 per row-step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuralsteganography_b200.coder import ArithmeticStreams
B, T, STEPS = int(os.environ.get("STREAMS", "2048")), 6, int(os.environ.get("STEPS", "60"))
rng = np.random.default_rng(7)
total = diff = 0
for V, precision, topk, temp, scale in ((50257, 26, 50257, 1.0, 3.0), (50257, 26, 300, 0.9, 3.0), (50257, 20, 40, 1.2, 2.0),
                                        (42001, 26, 512, 0.8, 4.0), (50257, 16, 5, 1.0, 1.0), (50257, 31, 50257, 1.0, 2.5)):
    g = torch.Generator(device="cuda").manual_seed(int(rng.integers(1 << 30)))
    pool = [torch.randn(B, V, generator=g, device="cuda") * scale for _ in range(T)]
    words = torch.from_numpy(rng.integers(0, 1 << 32, size=(B, 130), dtype=np.uint64).astype(np.uint32).view(np.int32))
    lens = torch.full((B,), 4096, dtype=torch.int32)
    a = ArithmeticStreams(B, V, precision=precision, temp=temp, topk=topk, token_cap=STEPS + 2)
    b = ArithmeticStreams(B, V, precision=precision, temp=temp, topk=topk, token_cap=STEPS + 2, force_exact=True)
    a.set_packed_messages(words, lens); b.set_packed_messages(words, lens)
    d_enc = 0
    for t in range(STEPS):
        a.encode_step(pool[t % T]); b.encode_step(pool[t % T])
        bad = (a.lo != b.lo) | (a.hi != b.hi) | (a.cursor != b.cursor) | (a.tokens[:, t] != b.tokens[:, t])
        n = int(bad.sum().item())
        if n:
            d_enc += n
            for name in ("lo", "hi", "cursor", "phase", "ntok"):
                getattr(a, name).copy_(getattr(b, name))
            a.tokens.copy_(b.tokens)
    handed = int(((a.status & 4) != 0).sum().item())
    toks, n = b.tokens.clone(), b.ntok.clone()
    da = ArithmeticStreams(B, V, precision=precision, temp=temp, topk=topk, token_cap=STEPS + 2)
    db = ArithmeticStreams(B, V, precision=precision, temp=temp, topk=topk, token_cap=STEPS + 2, force_exact=True)
    da.set_token_tensor(toks, n); db.set_token_tensor(toks, n)
    d_dec = 0
    for t in range(STEPS):
        da.decode_step(pool[t % T]); db.decode_step(pool[t % T])
        bad = (da.lo != db.lo) | (da.hi != db.hi) | (da.out_len != db.out_len)
        k = int(bad.sum().item())
        if k:
            d_dec += k
            for name in ("lo", "hi", "phase", "ntok", "out_len"):
                getattr(da, name).copy_(getattr(db, name))
            da.out_bits.copy_(db.out_bits)
    same_bits = bool((da.out_bits == db.out_bits).all().item()) if d_dec == 0 else None
    total += 2 * B * STEPS; diff += d_enc + d_dec
    print("V %d precision %d topk %d temp %.1f: %d row-steps each way, differences enc %d dec %d, rows ever handed over %d, decoded bits equal: %s"
          % (V, precision, topk, temp, B * STEPS, d_enc, d_dec, handed, same_bits))
print("total row-steps %d, differences %d" % (total, diff))
