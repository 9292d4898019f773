"""Config 2 end to end: GPT-2-small-shaped random-init trunk (PyTorch, cuBLAS) + arithmetic coder, whole loop
device-resident under one CUDA graph.  Reports tokens/s for B streams and the split trunk / coder."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuralsteganography_b200.generation import StegoGenerator
from neuralsteganography_b200.lm import random_init_model
import numpy as np

B = int(os.environ.get("STREAMS", "256"))
BITS = int(os.environ.get("BITS", "1024"))
dtype = {"fp32": torch.float32, "tf32": torch.float32, "bf16": torch.bfloat16}[os.environ.get("TRUNK", "fp32")]
tf32 = os.environ.get("TRUNK", "fp32") == "tf32"
_tok, model = random_init_model("gpt2")
model = model.cuda()
ctx = torch.tensor([50256, 11, 22])
msgs = [np.random.default_rng(10 + r).integers(0, 2, BITS).tolist() for r in range(B)]
for graph in (True, False):
    gen = StegoGenerator(model, B, max_len=512, precision=26, temp=0.9, topk=300, use_graph=graph, trunk_dtype=dtype, trunk_tf32=tf32)
    gen.encode(ctx, msgs, poll_every=32)            # warm-up (builds, captures)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    toks = gen.encode(ctx, msgs, poll_every=32)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = sum(len(t) for t in toks)
    print("%s: %d streams, %d message bits each: %d tokens in %.3f s = %.0f tok/s, %.2f ms per loop step (%d steps), %.2f bits/token"
          % ("CUDA graph" if graph else "eager     ", B, BITS, n, dt, n / dt, 1e3 * dt / gen.steps_run, gen.steps_run, B * BITS / n))
# split: trunk step alone, coder step alone (same shapes)
tok = torch.zeros(B, dtype=torch.long, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gen.trunk.reset(); gen._prefill(ctx)
for _ in range(5): gen.trunk.step(tok)
e0.record()
for _ in range(50): gen.trunk.step(tok)
e1.record(); torch.cuda.synchronize()
print("trunk step alone (eager, %s): %.3f ms" % (os.environ.get("TRUNK", "fp32"), e0.elapsed_time(e1) / 50))
from neuralsteganography_b200.coder import ArithmeticStreams
st = ArithmeticStreams(B, gen.V, precision=26, temp=0.9, topk=300, token_cap=128)
st.set_messages(msgs)
logits = torch.randn(B, gen.V, device="cuda") * 3.0
for _ in range(5): st.encode_step(logits)
e0.record()
for _ in range(50): st.encode_step(logits)
e1.record(); torch.cuda.synchronize()
print("coder step alone: %.3f ms" % (e0.elapsed_time(e1) / 50))
