"""Parity of the CUDA arithmetic coder (through the C ABI) with the oracle and the reference goldens."""
import numpy as np
import pytest
import torch

from oracle import ac_oracle as O
from oracle.inputs import logits_pool, make_distinct, message_bits, rows_for

pytestmark = pytest.mark.gpu


def _streams(*a, **k):
    from neuralsteganography_b200.coder import ArithmeticStreams
    return ArithmeticStreams(*a, **k)


def _run_encode_traced(st, fn, max_steps=400):
    """Step with a host sync per token so the per-step trace can be compared (tests only)."""
    trace = [[] for _ in range(st.B)]
    for t in range(max_steps):
        phase_before = st.phase.cpu().numpy().copy()
        st.encode_step(fn(t))
        torch.cuda.synchronize()
        tr = st.trace.cpu().numpy()
        nb = st.nbits.cpu().numpy()
        lo = st.lo.cpu().numpy(); hi = st.hi.cpu().numpy()
        for r in range(st.B):
            if phase_before[r] == 0:
                trace[r].append([int(tr[r, 0]), int(tr[r, 1]), int(nb[r]), int(lo[r]), int(hi[r])])
        if st.all_done():
            break
    return trace


@pytest.mark.parametrize("force_exact", [False, True])
def test_golden_cases_bit_exact(golden_dir, cases, force_exact):
    """Reference goldens through both kernels: throughput path (+ hand-over) and exact path only."""
    from gpu_util import PoolLogits, load_case
    for cfg in cases["ac"]:
        data, pool = load_case(golden_dir, cfg)
        S = cfg["streams"]
        fn = PoolLogits(pool, S)
        st = _streams(S, cfg["V"], precision=cfg["precision"], temp=cfg["temp"], topk=cfg["topk"],
                      token_cap=128, trace=True, force_exact=force_exact)
        st.set_messages([data["msg_%d" % s].tolist() for s in range(S)])
        trace = _run_encode_traced(st, fn)
        toks = st.token_lists()
        assert int((st.status & 11).sum().item()) == 0, cfg["name"]
        took_exact = bool((st.status & 4).any().item())
        if force_exact or cfg["precision"] > 31:
            assert not took_exact                                  # nothing was handed over: exact kernel only
        for s in range(S):
            assert toks[s] == data["tokens_%d" % s].tolist(), (cfg["name"], s, "tokens")
            want = data["trace_%d" % s][:, :5]                 # new_bottom, new_top, nbits, lo, hi
            assert np.array_equal(np.asarray(trace[s]), want), (cfg["name"], s, "interval trace")
        # decode the reference's tokens
        st.set_tokens([data["tokens_%d" % s].tolist() for s in range(S)])
        bits = st.decode(fn)
        assert int((st.status & 11).sum().item()) == 0
        for s in range(S):
            assert bits[s] == data["decoded_%d" % s].tolist(), (cfg["name"], s, "decoded bits")
            msg = data["msg_%d" % s].tolist()
            assert bits[s][: len(msg)] == msg


@pytest.mark.parametrize("precision,topk,temp", [(26, 50257, 1.0), (26, 300, 0.9), (16, 50000, 1.0)])
def test_integer_cdf_matches_oracle(precision, topk, temp):
    """Bin widths of every token (the integer CDF of code_base/arithmetic.py:146-158), mid-stream ranges."""
    V, B = 50257, 8
    pool = logits_pool(77, B, V, 3.0)
    st = _streams(B, V, precision=precision, temp=temp, topk=topk)
    rng = np.random.default_rng(3)
    los, his = [], []
    for r in range(B):
        width = int(rng.integers(2, 1 << precision)) if r else (1 << precision)
        lo = int(rng.integers(0, (1 << precision) - width + 1))
        los.append(lo); his.append(lo + width)
    st.lo.copy_(torch.tensor(los)); st.hi.copy_(torch.tensor(his))
    q, meta = st.debug_bins(torch.from_numpy(pool).cuda())
    q = q.cpu().numpy(); meta = meta.cpu().numpy()
    for r in range(B):
        masked = O.mask_row(pool[r], O.ENC_MASK)
        s, order = O.sort_desc(masked)
        cum, k = O.integer_cdf(O.softmax_f64(s, temp), his[r] - los[r], topk)
        widths = np.diff(np.concatenate([[0], cum]))
        got = q[r][order[:k]].copy()
        got[0] += meta[r, 1]                                   # slack goes to rank 0 (:158)
        assert np.array_equal(got, widths), (r, "bin widths")
        assert q[r][order[k:]].sum() == 0
        assert meta[r, 3] == his[r] - los[r]


def test_roundtrip_full_vocab_many_streams():
    """Config-3 shape property: encode -> decode recovers every message exactly.

    A stream whose interval collapses onto the midpoint (range 2) can stall on a cyclic logits pool --
    the reference loops there too -- so the loop is bounded and unfinished streams are checked on
    the bits they did consume.
    """
    V, B, nbits, P, STEPS = 50257, 512, 96, 3, 40
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1234)
    pool = [torch.randn(B, V, generator=g, device=dev) * 3.0 for _ in range(P)]
    fn = lambda t: pool[t % P]
    msgs = [message_bits(4321 + r, nbits - (r % 5)).tolist() for r in range(B)]
    st = _streams(B, V, precision=26, temp=1.0, topk=V, token_cap=64)
    st.set_messages(msgs)
    toks = st.encode(fn, poll_every=4, max_steps=STEPS)
    done = (st.phase.cpu().numpy() == 2)
    cursor = st.cursor.cpu().numpy()
    assert done.sum() >= B - 4
    assert int((st.status & 2).sum().item()) == 0
    st2 = _streams(B, V, precision=26, temp=1.0, topk=V, token_cap=64)
    st2.set_tokens(toks)
    bits = st2.decode(fn)
    for r in range(B):
        n = len(msgs[r]) if done[r] else int(cursor[r])
        assert bits[r][:n] == msgs[r][:n], r
    # a subset against the oracle on the same logits
    for r in (0, 17, 255, 320, 511):
        rows = lambda t, r=r: pool[t % P][r].cpu().numpy()
        res = O.encode_stream(rows, msgs[r], temp=1.0, precision=26, topk=V, max_steps=len(toks[r]))
        assert res.tokens == toks[r], r


def test_fast_and_exact_kernels_agree():
    """Same rows, same ranges: the throughput kernel and the exact kernel emit identical tokens/intervals."""
    V, B, T = 50257, 64, 5
    g = torch.Generator(device="cuda").manual_seed(99)
    pool = [torch.randn(B, V, generator=g, device="cuda") * (1.0 + 0.5 * p) for p in range(T)]
    fn = lambda t: pool[t % T]
    msgs = [message_bits(900 + r, 200).tolist() for r in range(B)]
    outs = []
    for force in (False, True):
        for temp in (1.0, 0.8):
            st = _streams(B, V, precision=26, temp=temp, topk=V, token_cap=40, trace=True, force_exact=force)
            st.set_messages(msgs)
            st.encode(fn, poll_every=8, max_steps=24)
            outs.append((force, temp, st.tokens.clone(), st.lo.clone(), st.hi.clone(), st.cursor.clone(),
                         int((st.status & 4).sum().item())))
    for temp in (1.0, 0.8):
        a = [o for o in outs if o[1] == temp and not o[0]][0]
        b = [o for o in outs if o[1] == temp and o[0]][0]
        # hand-overs only happen where the interval has collapsed (fewer than 2 tokens above 1/range)
        assert a[6] < 0.25 * B * 24
        for k in (2, 3, 4, 5):
            assert torch.equal(a[k], b[k]), (temp, k)


def test_ragged_empty_and_single_bit_messages():
    V, B = 2048, 5
    pool = logits_pool(11, 16, V, 3.0)
    from gpu_util import PoolLogits
    fn = PoolLogits(pool, B)
    msgs = [[], [1], [0, 1, 1], message_bits(1, 40).tolist(), message_bits(2, 200).tolist()]
    st = _streams(B, V, precision=16, temp=1.0, topk=50000, token_cap=128)
    st.set_messages(msgs)
    toks = st.encode(fn, poll_every=1)
    assert toks[0] == []
    for r in range(B):
        res = O.encode_stream(rows_for(pool, r), msgs[r], temp=1.0, precision=16, topk=50000, max_steps=128)
        assert toks[r] == res.tokens, r
    st.set_tokens(toks)
    bits = st.decode(fn)
    for r in range(B):
        want, _ = O.decode_stream(rows_for(pool, r), toks[r], temp=1.0, precision=16, topk=50000)
        assert bits[r] == want, r


def test_ties_follow_lower_id_first():
    """Rows with many exactly equal logits: order is (value desc, id asc) like the oracle."""
    V, B = 4096, 4
    rng = np.random.default_rng(5)
    pool = np.round(rng.standard_normal((8, V)).astype(np.float32) * 4) / 2      # heavy duplication
    from gpu_util import PoolLogits
    fn = PoolLogits(pool, B)
    msgs = [message_bits(50 + r, 120).tolist() for r in range(B)]
    st = _streams(B, V, precision=20, temp=1.0, topk=50000, token_cap=128)
    st.set_messages(msgs)
    toks = st.encode(fn, poll_every=2)
    flagged = (st.status.cpu().numpy() & 2) != 0
    for r in range(B):
        if flagged[r]:
            continue                                            # bucket overflow is reported, not silent
        res = O.encode_stream(rows_for(pool, r), msgs[r], temp=1.0, precision=20, topk=50000, max_steps=128)
        assert toks[r] == res.tokens, r


def test_finish_sent_tail_and_out_of_range_token():
    V, B = 2048, 2
    pool = logits_pool(21, 8, V, 3.0)
    from gpu_util import PoolLogits
    fn = PoolLogits(pool, B)
    sent_end = torch.zeros(V, dtype=torch.uint8, device="cuda")
    top_per_step = [int(np.argmax(O.mask_row(rows_for(pool, 0)(t), O.ENC_MASK))) for t in range(40)]
    msgs = [message_bits(9, 24).tolist(), message_bits(10, 24).tolist()]
    st = _streams(B, V, precision=16, temp=1.0, topk=50000, token_cap=64, finish_sent=True, sent_end=sent_end)
    st.set_messages(msgs)
    base = O.encode_stream(rows_for(pool, 0), msgs[0], temp=1.0, precision=16, topk=50000, max_steps=64).tokens
    stop_at = len(base) + 2
    sent_end[top_per_step[stop_at]] = 1                         # third tail token ends the sentence
    toks = st.encode(fn, poll_every=1, max_steps=40)
    assert toks[0][: len(base)] == base
    tail = toks[0][len(base):]
    assert tail == top_per_step[len(base): len(base) + len(tail)]      # rank-0 tokens (arithmetic.py:135-137)
    assert tail[-1] == top_per_step[stop_at] or len(toks[0]) == 40
    # decode with a token the kept set cannot contain (masked id V-1): flagged, coded as rank 0 (:342)
    bad = [list(toks[0][: len(base)]), list(toks[1])]
    bad[0][1] = V - 1
    st.finish_sent = False
    st.set_tokens(bad)
    st.decode(fn)
    assert int(st.status[0].item()) & 1


def test_rejects_cpu_tensors_and_bad_shapes():
    from neuralsteganography_b200 import NativeLibraryError
    st = _streams(2, 2048, precision=16)
    st.set_messages([[1, 0], [0, 1]])
    with pytest.raises(NativeLibraryError):
        st.encode_step(torch.zeros(2, 2048))
    with pytest.raises(NativeLibraryError):
        st.encode_step(torch.zeros(3, 2048, device="cuda"))


def test_device_statistics_match_oracle():
    """a6: log p(selected), KL(q_hat||p) and entropy per step (code_base/arithmetic.py:192-198) vs the oracle."""
    V, B = 50257, 3
    pool = logits_pool(123, 6, V, 3.0)
    from gpu_util import PoolLogits
    fn = PoolLogits(pool, B)
    for temp, topk, precision in ((0.9, 300, 26), (1.0, V, 26)):
        msgs = [message_bits(60 + r, 64).tolist() for r in range(B)]
        st = _streams(B, V, precision=precision, temp=temp, topk=topk, token_cap=32)
        st.stats = torch.zeros(B, 3, dtype=torch.float64, device="cuda")
        st.set_messages(msgs)
        lo = [0] * B; hi = [1 << precision] * B; cur = [0] * B
        for t in range(4):
            st.encode_step(fn(t))
            torch.cuda.synchronize()
            got = st.stats.cpu().numpy()
            for r in range(B):
                tr, ncur, stats, _ = O.encode_step(rows_for(pool, r)(t), lo[r], hi[r], msgs[r], cur[r], temp=temp,
                                                   precision=precision, topk=topk, want_stats=True)
                assert int(st.tokens[r, t]) == tr.token
                np.testing.assert_allclose(got[r], stats, rtol=1e-9, atol=1e-12)
                lo[r], hi[r], cur[r] = tr.lo, tr.hi, ncur


@pytest.mark.parametrize("V,precision,topk,temp,scale,quant", [
    (50257, 26, 300, 0.9, 3.0, 0.0),      # config 2's coder settings
    (50257, 26, 2, 1.0, 3.0, 0.0),        # smallest top-k
    (50257, 16, 17, 0.7, 3.0, 0.0),       # low precision: overfill in the rank form
    (50257, 12, 512, 1.3, 1.0, 0.0),      # flat rows, largest top-k of the path, heavy overfill
    (42001, 26, 300, 1.0, 8.0, 0.0),      # peaked rows: often fewer than topk tokens above the cutoff
    (50257, 26, 64, 1.0, 3.0, 0.25),      # quantised logits: many exact ties across the top-k boundary
    (5000, 20, 100, 1.0, 2.0, 0.5),       # small vocabulary, dense ties (boundary bucket overflow -> hand-over)
    (42001, 26, 384, 0.8, 4.0, 0.0),      # a top-k between the usual one and the largest of the path
    (9000, 20, 10, 1.0, 2.0, 0.0),        # smallest vocabularies of the sweep kernel (three rounds of the sample selection)
    (50257, 16, 300, 0.9, 6.0, 0.0),      # top-k rarely binds at this range: the sweep kernel leaves the rows to the threshold form
])
@pytest.mark.parametrize("variant", [0, 2])
def test_rank_form_fast_path_matches_exact_kernel(V, precision, topk, temp, scale, quant, variant):
    """The top-k-binding paths of the throughput kernels (variant 0: the sweep kernel ns_topk.cuh + the row-resident
    kernel on what it leaves; variant 2: the row-resident kernel alone) against the exact kernel, encode and decode:
    same tokens, intervals, cursors, recovered bits; plus the oracle on a few streams."""
    B, T, steps = 48, 4, 20
    g = torch.Generator(device="cuda").manual_seed(1000 + topk)
    pool = [torch.randn(B, V, generator=g, device="cuda") * scale for _ in range(T)]
    if quant:
        pool = [(p / quant).round() * quant for p in pool]
    fn = lambda t: pool[t % T]
    msgs = [message_bits(1500 + r, 256).tolist() for r in range(B)]
    enc = {}
    for force in (False, True):
        st = _streams(B, V, precision=precision, temp=temp, topk=topk, token_cap=steps + 2, trace=True, force_exact=force,
                      variant=variant)
        st.set_messages(msgs)
        st.encode(fn, poll_every=64, max_steps=steps)
        enc[force] = st
    a, b = enc[False], enc[True]
    for name in ("tokens", "lo", "hi", "cursor", "ntok"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert int((a.status & 3).sum().item()) == 0
    if (topk, quant, scale) == (300, 0.0, 3.0):
        # config 2's settings on ordinary rows: the rank-form path of the throughput kernel codes every row itself
        assert int(((a.status & 4) != 0).sum().item()) == 0
        # ... and with variant 0 the sweep kernel does, leaving nothing to the row-resident one
        assert int(((a.status & 16) != 0).sum().item()) == 0
    if variant == 2:
        assert int(((a.status & 16) != 0).sum().item()) == 0
    assert int(a.rank_ws[:2].abs().sum().item()) == 0          # the work list is empty again after every step
    toks = a.token_lists()
    dec = {}
    for force in (False, True):
        st = _streams(B, V, precision=precision, temp=temp, topk=topk, token_cap=steps + 2, force_exact=force, variant=variant)
        st.set_tokens(toks)
        dec[force] = st.decode(fn)
    assert dec[False] == dec[True]
    cur = a.cursor.cpu().numpy()
    for r in range(B):
        n = int(min(cur[r], len(msgs[r])))
        assert dec[False][r][:n] == msgs[r][:n], r
    for r in (0, B - 1):
        rows = lambda t, r=r: pool[t % T][r].cpu().numpy()
        res = O.encode_stream(rows, msgs[r], temp=temp, precision=precision, topk=topk, max_steps=steps)
        assert toks[r][: len(res.tokens)] == res.tokens[: len(toks[r])], r


def test_rank_form_decode_of_foreign_token():
    """A cover token outside the top-k: flagged, coded as rank 0 (arithmetic.py:342), same as the exact kernel."""
    V, B = 50257, 8
    g = torch.Generator(device="cuda").manual_seed(77)
    logits = torch.randn(B, V, generator=g, device="cuda") * 3.0
    worst = logits.argmin(dim=1).cpu().tolist()
    outs = []
    for force in (False, True):
        st = _streams(B, V, precision=26, temp=0.9, topk=300, token_cap=4, force_exact=force)
        st.set_tokens([[w, w] for w in worst])
        st.decode_step(logits)
        outs.append((st.lo.clone(), st.hi.clone(), st.out_len.clone(), st.out_bits.clone(), (st.status & 1).clone()))
    for x, y in zip(*outs):
        assert torch.equal(x, y)
    assert bool((outs[0][4] == 1).all().item())


def test_sweep_kernel_tail_and_masks_match_exact_kernel():
    """finish_sent tails and forbidden tokens through the rank form's sweep kernel (variant 0): the forbidden ids carry
    the largest logits of every row (they must get probability 0), short messages put the streams into the tail after
    a few tokens (those rows are left to the row-resident kernel)."""
    V, B, steps = 50257, 16, 14
    g = torch.Generator(device="cuda").manual_seed(99)
    pool = [torch.randn(B, V, generator=g, device="cuda") * 3.0 for _ in range(3)]
    for p in pool:
        p[:, V - 1] = 40.0
        p[:, 628] = 35.0
    sent_end = (torch.rand(V, generator=g, device="cuda") < 0.3).to(torch.uint8)
    msgs = [message_bits(300 + r, 30 + r).tolist() for r in range(B)]
    out = {}
    for force in (False, True):
        st = _streams(B, V, precision=26, temp=0.9, topk=300, token_cap=steps + 2, finish_sent=True, sent_end=sent_end,
                      force_exact=force, variant=0)
        st.set_messages(msgs)
        st.encode(lambda t: pool[t % 3], poll_every=64, max_steps=steps)
        out[force] = st
    a, b = out[False], out[True]
    for name in ("tokens", "lo", "hi", "cursor", "ntok", "phase"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert int((a.status & 3).sum().item()) == 0
    assert bool(((a.status & 16) != 0).any().item())            # the tails went through the work list
    assert not bool((a.tokens == V - 1).any().item()) and not bool((a.tokens == 628).any().item())
    assert int(a.rank_ws[:2].abs().sum().item()) == 0
