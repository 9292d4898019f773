"""Print why rows were handed to the exact kernel (status bits 8..15), for a few input families."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from neuralsteganography_b200.coder import ArithmeticStreams
from oracle.inputs import message_bits
V, B, T = 50257, 64, 5
g = torch.Generator(device="cuda").manual_seed(99)
for scale_fn, name in ((lambda p: 1.0 + 0.5 * p, "scales 1..3"), (lambda p: 3.0, "scale 3")):
    pool = [torch.randn(B, V, generator=g, device="cuda") * scale_fn(p) for p in range(T)]
    msgs = [message_bits(900 + r, 200).tolist() for r in range(B)]
    for temp in (1.0, 0.8):
        st = ArithmeticStreams(B, V, precision=26, temp=temp, topk=V, token_cap=40, trace=True)
        st.set_messages(msgs)
        hist = {}
        for t in range(24):
            st.status.zero_()
            lo0 = st.lo.cpu().numpy().copy(); hi0 = st.hi.cpu().numpy().copy(); ph0 = st.phase.cpu().numpy().copy()
            st.encode_step(pool[t % T])
            torch.cuda.synchronize()
            s = st.status.cpu().numpy()
            R = (st.trace[:, 1] - st.trace[:, 0]).cpu().numpy()
            for r in np.nonzero(s & 4)[0]:
                why = int(s[r]) >> 8
                hist[why] = hist.get(why, 0) + 1
                if hist[why] <= 2:
                    print("   row", r, "step", t, "why", why, "status", hex(int(s[r])), "R", int(hi0[r] - lo0[r]), "lo", int(lo0[r]), "phase", int(ph0[r]),
                          "cursor", int(st.cursor[r]), "len", len(msgs[r]))
        print(name, "temp", temp, "hand-over reasons {why: count}:", hist)
