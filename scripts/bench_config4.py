"""BASELINE config 4: gpt2-fa-shaped random-init trunk (42001 tokens), 1024 streams: cover tokens -> messages.
Times the step-wise decode loop (CUDA graphs) against the teacher-forced tiled decode; prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuralsteganography_b200.generation import StegoGenerator
from neuralsteganography_b200.lm import random_init_model

B = int(os.environ.get("STREAMS", "1024")); NBITS = int(os.environ.get("BITS", "512")); TILE = int(os.environ.get("TILE", "32"))
_tok, model = random_init_model("gpt2-fa")
model = model.cuda()
gen = StegoGenerator(model, B, max_len=256, precision=26, temp=0.9, topk=300, use_graph=True)
ctx = torch.tensor([5, 11, 22])
rng = np.random.default_rng(44)
msgs = [rng.integers(0, 2, NBITS).tolist() for r in range(B)]
toks = gen.encode(ctx, msgs, poll_every=16)
ntok = sum(len(t) for t in toks)
def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out
t_seq, seq = timed(lambda: gen.decode(ctx, toks, poll_every=64))
t_til, til = timed(lambda: gen.decode_prefill(ctx, toks, tile=TILE))
ok = all(a[:NBITS] == m and b[:NBITS] == m for a, b, m in zip(seq, til, msgs))
print(json.dumps({"workload": "config 4: %d streams, V=42001, %d cover tokens, temp 0.9 / precision 26 / topk 300, fp32 trunk" % (B, ntok),
                  "stepwise_decode_tokens_per_sec": ntok / t_seq, "tiled_decode_tokens_per_sec": ntok / t_til,
                  "tiled_over_stepwise": t_seq / t_til, "all_messages_recovered": ok, "tile": TILE}))
