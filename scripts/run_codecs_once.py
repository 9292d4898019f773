"""A few steps of each comparison codec at V = 50257 (profiling target; STREAMS rows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuralsteganography_b200.codecs import CodecStreams
V, B = 50257, int(os.environ.get("STREAMS", "592"))
g = torch.Generator(device="cuda").manual_seed(1234)
pool = [torch.randn(B, V, generator=g, device="cuda") * 3.0 for _ in range(2)]
rng = np.random.default_rng(0)
words = torch.from_numpy(rng.integers(0, 1 << 32, size=(B, 130), dtype=np.uint64).astype(np.uint32).view(np.int32))
lens = torch.full((B,), 4096, dtype=torch.int32)
for kind, kw in (("bins", dict(param=3)), ("rank", {}), ("huffman", dict(param=3))):
    cs = CodecStreams(kind, B, V, token_cap=16, **kw)
    cs.set_packed_messages(words, lens)
    for t in range(4):
        cs.encode_step(pool[t % 2])
    torch.cuda.synchronize()
    print(kind, "ok", int(cs.cursor.sum().item()))
