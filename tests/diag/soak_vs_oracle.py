"""the kernels against the CPU oracle on a 64-stream x 64-step subset of the headline shape,
for the full distribution (config 3) and for temp 0.9 / topk 300 (config 2's coder settings); encode tokens and
decoded bits.  The oracle runs one process per stream group on the host cores."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import multiprocessing as mp
import numpy as np

V, B, T, STEPS = 50257, int(os.environ.get("STREAMS", "64")), 8, int(os.environ.get("STEPS", "64"))


def oracle_stream(args):
    rows, msg, kw, steps = args
    from oracle import ac_oracle as O
    res = O.encode_stream(lambda t: rows[t % len(rows)], msg, max_steps=steps, **kw)
    bits, _ = O.decode_stream(lambda t: rows[t % len(rows)], res.tokens, **kw)
    return res.tokens, bits


def main():
    import torch
    from neuralsteganography_b200.coder import ArithmeticStreams
    rng = np.random.default_rng(5)
    total = bad = 0
    for kw in (dict(temp=1.0, precision=26, topk=V), dict(temp=0.9, precision=26, topk=300)):
        g = torch.Generator(device="cuda").manual_seed(int(rng.integers(1 << 30)))
        pool = [torch.randn(B, V, generator=g, device="cuda") * 3.0 for _ in range(T)]
        msgs = [rng.integers(0, 2, 4096).tolist() for _ in range(B)]
        st = ArithmeticStreams(B, V, token_cap=STEPS + 2, **kw)
        st.set_messages(msgs)
        for t in range(STEPS):
            st.encode_step(pool[t % T])
        toks = st.token_lists()
        st.set_tokens(toks)
        bits = st.decode(lambda t: pool[t % T])
        host = [p.cpu().numpy() for p in pool]
        jobs = [([host[t][r] for t in range(T)], msgs[r], kw, STEPS) for r in range(B)]
        with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, 16)) as pl:
            ref = pl.map(oracle_stream, jobs)
        d = sum(1 for r in range(B) if toks[r] != ref[r][0] or bits[r] != ref[r][1])
        total += B * STEPS; bad += d
        print("%s: %d streams x %d steps vs the oracle: streams differing %d" % (kw, B, STEPS, d))
    print("row-steps %d, streams differing %d" % (total, bad))


if __name__ == "__main__":
    main()
