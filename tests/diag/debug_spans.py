import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stub_tokenizer import StubTokenizer
from neuralsteganography_b200 import framing as F
from neuralsteganography_b200.lm import B200ArithmeticLM
from transformers import GPT2Config, GPT2LMHeadModel
torch.manual_seed(7)
model = GPT2LMHeadModel(GPT2Config(n_layer=2, n_embd=64, n_head=2, vocab_size=2048)).eval().cuda()
tok = StubTokenizer()
lm = B200ArithmeticLM(model, tok, max_len=1024)
q = {"temp": 1.0, "precision": 16, "topk": 6, "finish_sent": True}
seed_text = "helloworld"
res = F.stego_encode(b"meet at dawn", chunk_bytes=6, use_crc=True, ecc="none", quality=q, seed_text=seed_text, lm=lm, msg_id="t")
spans = [list(s) for s in res]
print("span lens", [len(s) for s in spans])
text = tok.decode(lm.encode_seed(seed_text) + [t for s in spans for t in s])
got = lm.text_to_spans(text, seed_text, quality=q)
print("got lens", [len(s) for s in got])
for a, b in zip(spans, got):
    n = min(len(a), len(b))
    d = next((i for i in range(n) if a[i] != b[i]), n)
    print("first diff at", d, "of", len(a), len(b))
    print(" want", a[max(0, d - 3): d + 6], [tok.decode([t]) for t in a[max(0, d - 3): d + 6]])
    print(" got ", b[max(0, d - 3): d + 6], [tok.decode([t]) for t in b[max(0, d - 3): d + 6]])
    print(" tails", [tok.decode([t]) for t in a[-4:]], [tok.decode([t]) for t in b[-4:]])
