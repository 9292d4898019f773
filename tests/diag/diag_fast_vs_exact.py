"""Lock-step comparison of the throughput kernel and the exact kernel; on the first difference print
both traces and the oracle's answer for that row."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from neuralsteganography_b200.coder import ArithmeticStreams
from oracle import ac_oracle as O
from oracle.inputs import message_bits
V, B, T = 50257, 64, 5
g = torch.Generator(device="cuda").manual_seed(99)
pool = [torch.randn(B, V, generator=g, device="cuda") * (1.0 + 0.5 * p) for p in range(T)]
msgs = [message_bits(900 + r, 200).tolist() for r in range(B)]
temp = 1.0
a = ArithmeticStreams(B, V, precision=26, temp=temp, topk=V, token_cap=40, trace=True)
b = ArithmeticStreams(B, V, precision=26, temp=temp, topk=V, token_cap=40, trace=True, force_exact=True)
a.set_messages(msgs); b.set_messages(msgs)
shown = 0
for t in range(24):
    lo0 = a.lo.cpu().numpy().copy(); hi0 = a.hi.cpu().numpy().copy(); cur0 = a.cursor.cpu().numpy().copy()
    ph0 = a.phase.cpu().numpy().copy()
    a.status.zero_()
    a.encode_step(pool[t % T]); b.encode_step(pool[t % T]); torch.cuda.synchronize()
    ta = a.trace.cpu().numpy(); tb = b.trace.cpu().numpy()
    tka = a.tokens[:, t].cpu().numpy(); tkb = b.tokens[:, t].cpu().numpy()
    sa = a.status.cpu().numpy()
    bad = np.nonzero((tka != tkb) | (ta[:, 0] != tb[:, 0]) | (ta[:, 1] != tb[:, 1]))[0]
    bad = [r for r in bad if ph0[r] == 0]
    for r in bad[:3]:
        if shown >= 4: break
        shown += 1
        R = int(hi0[r] - lo0[r])
        print("step", t, "row", r, "R", R, "lo", int(lo0[r]), "status_fast", hex(int(sa[r])))
        print("  fast : tok", tka[r], "nb,nt,k,Q", ta[r].tolist())
        print("  exact: tok", tkb[r], "nb,nt,k,Q", tb[r].tolist())
        row = pool[t % T][r].cpu().numpy()
        st, _, _, cum = O.encode_step(row, int(lo0[r]), int(hi0[r]), msgs[r], int(cur0[r]), temp=temp, precision=26, topk=V)
        print("  oracle: tok", st.token, "nb,nt", st.new_bottom, st.new_top, "k", st.k, "sel", st.selection)
    if bad:
        # resync the fast stream to the exact one so later steps stay comparable
        for name in ("lo", "hi", "cursor", "phase", "ntok"):
            getattr(a, name).copy_(getattr(b, name))
        a.tokens.copy_(b.tokens)
print("done; mismatching row-steps shown:", shown)
