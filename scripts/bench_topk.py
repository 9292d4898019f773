"""Coder-only throughput at the reference's usual quality (temp 0.9, precision 26, topk 300: config 2's coder
settings) at the headline shape, with the share of rows the throughput kernel hands to the exact kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuralsteganography_b200.coder import ArithmeticStreams
V, B, P = 50257, int(os.environ.get("STREAMS", "4096")), 4
TOPK, TEMP = int(os.environ.get("TOPK", "300")), float(os.environ.get("TEMP", "0.9"))
g = torch.Generator(device="cuda").manual_seed(1234)
pool = [torch.randn(B, V, generator=g, device="cuda") * 3.0 for _ in range(P)]
rng = np.random.default_rng(0)
words = rng.integers(0, 1 << 32, size=(B, 130), dtype=np.uint64).astype(np.uint32)
for force in (False, True):
    st = ArithmeticStreams(B, V, precision=26, temp=TEMP, topk=TOPK, token_cap=64, force_exact=force)
    st.set_packed_messages(torch.from_numpy(words.view(np.int32)), torch.full((B,), 4096, dtype=torch.int32))
    for t in range(3): st.encode_step(pool[t % P])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20 if not force else 5
    c0 = int(st.cursor.sum().item())
    e0.record()
    for t in range(n): st.encode_step(pool[(3 + t) % P])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    handed = int(((st.status & 4) != 0).sum().item())
    deferred = int(((st.status & 16) != 0).sum().item())
    print("topk %d temp %.2f %s: %.3f ms/step  %.2f M tok/s  %.2f bits/token  rows ever handed to the exact kernel: %d of %d, ever left by the sweep kernel to the row-resident one: %d"
          % (TOPK, TEMP, "exact kernel only" if force else "throughput kernel", ms, B / ms / 1e3,
             (int(st.cursor.sum().item()) - c0) / (B * n), handed, B, deferred))
    if not force: toks_fast = st.tokens.clone()
    else: print("   same tokens as the throughput kernel on the common steps:", bool((toks_fast[:, :8] == st.tokens[:, :8]).all().item()))
