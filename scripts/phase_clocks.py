"""Per-phase SM-clock cycles of the throughput kernel at the bench configuration (thread 0 of every CTA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuralsteganography_b200.coder import ArithmeticStreams
V, B, P = 50257, int(os.environ.get("STREAMS", "4096")), 2
TOPK, TEMP = int(os.environ.get("TOPK", str(V))), float(os.environ.get("TEMP", "1.0"))
g = torch.Generator(device="cuda").manual_seed(1234)
pool = [torch.randn(B, V, generator=g, device="cuda") * 3.0 for _ in range(P)]
rng = np.random.default_rng(0)
words = rng.integers(0, 1 << 32, size=(B, 130), dtype=np.uint64).astype(np.uint32)
names = ["prologue", "L wait+estimate", "reduce+consts", "P1 exp pass", "FIX", "P2 q pass", "fixups+scan", "overfill sel",
         "target sel", "epilogue", "row-top barrier", "  rank: sample histogram", "  rank: scan + list candidates", "  rank: level-2 histogram + gather",
         "  rank: boundary+order+exp"]
for mode in ("enc", "dec"):
    st = ArithmeticStreams(B, V, precision=26, temp=TEMP, topk=TOPK, token_cap=32)
    st.set_packed_messages(torch.from_numpy(words.view(np.int32)), torch.full((B,), 4096, dtype=torch.int32))
    for t in range(3): st.encode_step(pool[t % P])
    if mode == "dec":
        toks, n = st.tokens.clone(), st.ntok.clone()
        st = ArithmeticStreams(B, V, precision=26, temp=TEMP, topk=TOPK, token_cap=32)
        st.set_token_tensor(toks, n)
    st.prof = torch.zeros(32, dtype=torch.int64, device="cuda")
    step = st.encode_step if mode == "enc" else st.decode_step
    for t in range(3): step(pool[t % P])
    torch.cuda.synchronize()
    pr = st.prof.cpu().numpy().astype(np.float64)
    if TOPK < V and os.environ.get("NS_AC_VARIANT", "0") != "2":
        rnames = ["sweep 1 (HBM)", "reductions + bound", "sweep 2 (L2)", "rank-form check", "group listed keys", "order + exp + widths", "overfill + search + update"]
        print(mode, "rank-form kernel: rows", int(pr[15]), "cycles/row (thread 0) %.0f" % (pr[:7].sum() / pr[15]))
        for k, nm in enumerate(rnames):
            print("   %-28s %8.0f" % (nm, pr[k] / pr[15]))
        continue
    rows = pr[15]
    if TOPK >= V:      # lean kernel: slots 11..14 are the steps inside the selection (not part of the sum above them)
        names[11:15] = ["  sel: locate bucket", "  sel: gather sweep", "  sel: gather barrier", "  sel: order candidates"]
    if pr[11:15].any():
        print("   (slots 11-14 are timed separately: a row is the sum of all lines below)")
    print(mode, "rows", int(rows), "cycles/row total %.0f" % (pr[:15].sum() / rows))
    if pr[16:25].any():
        print("   piece arrival (cycles after row start, as seen by warp 1):", " ".join("%.0f" % (x / rows) for x in pr[16:25]))
    if pr[25:29].any():
        print("   thread 0: entry %.0f, fence %.0f, expect_tx %.0f, bulk issue %.0f" % tuple(pr[25:29] / rows))
    for k, nm in enumerate(names):
        if k >= 11 and pr[k] == 0:
            continue
        print("   %-18s %8.0f cyc/row %5.1f%%" % (nm, pr[k] / rows, 100 * pr[k] / pr[:15].sum()))
