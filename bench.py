#!/usr/bin/env python
"""Benchmark of the coder step at BASELINE.json's headline configuration.

Workload (configs[2]): coder-only batch, 4096 streams x 50257 synthetic fp32 logits per GPU,
full distribution (topk = V), precision 26, temp 1.0.  One "step" = one arithmetic-coder encode
step over all streams of the rank (one kernel launch).  Inputs rotate over a pool of 4 logits
tensors (4 x 823 MB >> 126 MB L2), so every step streams its rows from HBM.

    python bench.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference      # the CPU codec (oracle port) on the host cores
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V = 50257
STREAMS = 4096
POOL = 4
PRECISION = 26
TEMP = 1.0
MSG_BITS = 4096
ALGO_BYTES_PER_TOKEN = 4 * V + 32          # SURVEY.md 8d / DESIGN.md
# dram__bytes_read.sum + dram__bytes_write.sum of ac_lean_kernel<1, 0, 0> from one `ncu --set full` capture
# (592 rows: 119.18 MB + 3.87 MB, profiles/r2_lean_ncu_summary.txt), per row
NCU_DRAM_BYTES_PER_TOKEN = (119176192 + 3870208) / 592
NCU_TRAFFIC_SOURCE = "profiles/r2_lean_ncu_summary.txt: ncu --set full capture of ac_lean_kernel<1,0,0> on 592 rows, scaled per row"
METRIC = "coder_tokens_per_sec"
WORKLOAD = "configs[2]: coder-only batch, 4096 streams x 50257 fp32 logits per GPU, full distribution, precision 26, temp 1.0"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# CPU codec (oracle port) -- baseline leg and --impl reference arm
# ------------------------------------------------------------------------------------------
_ROWS = None


def _cpu_rows(seed):
    global _ROWS
    if _ROWS is None or _ROWS[0] != seed:
        rng = np.random.Generator(np.random.PCG64(seed))
        _ROWS = (seed, [rng.standard_normal(V, dtype=np.float32) * np.float32(3.0) for _ in range(4)],
                 rng.integers(0, 2, MSG_BITS).tolist())
    return _ROWS[1], _ROWS[2]


def _cpu_worker(args):
    """One stream of the CPU codec: `steps` tokens of leg `kind`; returns (tokens, bits, seconds)."""
    seed, steps, kind = args
    from oracle import ac_oracle as O
    from oracle import codecs_oracle as K
    rows, msg = _cpu_rows(seed)
    fn = lambda t: rows[t % 4]
    t0 = time.perf_counter()
    if kind == "ac_encode_full":
        res = O.encode_stream(fn, msg, temp=TEMP, precision=PRECISION, topk=V, max_steps=steps, keep_trace=False)
        n, bits = len(res.tokens), res.bits_consumed
    elif kind == "ac_encode_topk300":
        res = O.encode_stream(fn, msg, temp=0.9, precision=PRECISION, topk=300, max_steps=steps, keep_trace=False)
        n, bits = len(res.tokens), res.bits_consumed
    elif kind == "ac_decode_full":
        res = O.encode_stream(fn, msg, temp=TEMP, precision=PRECISION, topk=V, max_steps=steps, keep_trace=False)
        t0 = time.perf_counter()                              # only the decode is timed
        out, _ = O.decode_stream(fn, res.tokens, temp=TEMP, precision=PRECISION, topk=V)
        n, bits = len(res.tokens), len(out)
    elif kind == "huffman_b3":
        toks, used = K.huffman_encode(fn, msg[: 3 * steps], 3)
        n, bits = len(toks), used
    elif kind == "bins_b3":
        toks = K.bins_encode(fn, msg[: 3 * steps], 3, V)
        toks = toks[0] if isinstance(toks, tuple) else toks
        n, bits = len(toks), 3 * len(toks)
    elif kind == "rank":
        payload = bytes(np.packbits(np.asarray(msg[: 15 * steps + 8], dtype=np.uint8))[: (15 * steps) // 8])
        toks, hist, total = K.rank_encode(fn, payload)
        n, bits = len(toks), total
    else:
        raise ValueError(kind)
    return n, bits, time.perf_counter() - t0


class CpuCodec:
    """All host cores, one independent stream per process (the reference codes one stream per call)."""

    def __init__(self, procs):
        import multiprocessing as mp
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        self.procs = procs
        self.pool = mp.get_context("fork").Pool(procs)

    def run(self, steps_per_stream, kind="ac_encode_full"):
        """-> (tokens/s, bits/s, tokens, wall seconds) of one bounded sample: `procs` streams x `steps_per_stream` tokens."""
        t0 = time.perf_counter()
        out = self.pool.map(_cpu_worker, [(100 + i, steps_per_stream, kind) for i in range(self.procs)])
        wall = time.perf_counter() - t0
        toks = sum(o[0] for o in out)
        bits = sum(o[1] for o in out)
        busy = max(o[2] for o in out)
        return toks / busy, bits / busy, toks, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_codec_throughput(steps_per_stream: int, procs: int):
    c = CpuCodec(procs)
    try:
        return c.run(steps_per_stream)
    finally:
        c.close()


def config_dict(B, strong=False):
    return {"workload": WORKLOAD, "streams_per_gpu": B, "vocab": V, "precision": PRECISION, "temp": TEMP,
            "topk": V, "l2": "inputs larger than L2: 4-entry logits pool, 823 MB per step",
            "parallelism": "streams sharded over ranks, no collective in the loop"}


def run_reference_arm(args):
    """The reference's CPU codec on the box's host cores.  One step = one bounded sample of the workload: every core
    codes `per` tokens of its own stream (the full 4096-stream step would take ~17 s per step on 16 cores);
    `ms_per_step` is the measured wall time of such a step, `value` the tokens/s it amounts to."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per = 32
    codec = CpuCodec(cores)
    for _ in range(max(1, min(args.warmup, 2))):
        codec.run(2)
    steps = max(1, min(args.steps, 60))
    toks = 0
    bits = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        tps_i, bps_i, n_i, wall_i = codec.run(per)
        toks += n_i
        bits += bps_i * (n_i / tps_i)
    wall = time.perf_counter() - t0
    codec.close()
    tps = toks / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(config_dict(STREAMS), sample="one step = %d streams x %d encode steps (one process per core), a bounded "
                       "sample of the 4096-stream step" % (cores, per),
                       note="CPU codec = oracle port of code_base/arithmetic.py (the Python reference is not present on the GPU box)"),
        "message_bits_per_sec": bits / wall,
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": "%d steps of %d independent streams x %d encode steps, V=50257, one process per core" % (steps, cores, per)},
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from neuralsteganography_b200.coder import ArithmeticStreams, pack_bits

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # CPU baseline first: its worker processes are forked before this process owns a CUDA context
    cores = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        codec = CpuCodec(cores)
        ctps, cbps, ctoks, cwall = codec.run(args.cpu_steps)
        legs = {}
        for kind, n in (("ac_decode_full", args.cpu_steps // 4), ("ac_encode_topk300", args.cpu_steps // 4),
                        ("huffman_b3", args.cpu_steps // 4), ("bins_b3", args.cpu_steps // 8), ("rank", args.cpu_steps // 8)):
            ltps, lbps, ltoks, lwall = codec.run(max(4, n), kind)
            legs[kind] = {"tokens_per_sec": ltps, "bits_per_sec": lbps, "tokens": ltoks}
        codec.close()
        cpu = {"value": ctps, "unit": "tokens/s", "cores": cores, "kind": "port",
               "message_bits_per_sec": cbps,
               "sample": "%d independent streams x %d encode steps of the same workload (V=50257, precision 26, "
                         "full distribution), one process per core; %d tokens in %.1f s wall" % (cores, args.cpu_steps, ctoks, cwall),
               "legs": legs}
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the coder has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weak scaling (the contract's default): --streams per GPU; --strong: --streams in total, split over the ranks
    B = args.streams if not args.strong else max(1, args.streams // world)
    K, W = args.steps, max(3, args.warmup)
    # synthetic inputs: pool of logits, random messages
    pool = []
    for p in range(POOL):
        g = torch.Generator(device=dev).manual_seed(1234 + 16 * rank + p)
        pool.append(torch.randn(B, V, generator=g, device=dev, dtype=torch.float32) * 3.0)
    rng = np.random.Generator(np.random.PCG64(4321 + rank))
    nwords = MSG_BITS // 32 + 2
    words = rng.integers(0, 1 << 32, size=(B, nwords), dtype=np.uint64).astype(np.uint32)
    lens = np.full(B, MSG_BITS, dtype=np.int32)
    cap = K + W + 8

    st = ArithmeticStreams(B, V, precision=PRECISION, temp=TEMP, topk=V, token_cap=cap, device=dev)
    st.set_packed_messages(torch.from_numpy(words.view(np.int32)), torch.from_numpy(lens))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(W):
        st.encode_step(pool[t % POOL])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur0 = int(st.cursor.sum().item())
    barrier()
    ev0.record()
    for t in range(K):
        st.encode_step(pool[(W + t) % POOL])
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    bits = int(st.cursor.sum().item()) - cur0
    live = int((st.phase == 0).sum().item())

    # decode the same cover tokens (not part of `value`; reported beside it)
    toks = st.tokens.clone()
    ntok = st.ntok.clone()
    dec = ArithmeticStreams(B, V, precision=PRECISION, temp=TEMP, topk=V, token_cap=cap, device=dev)
    dec.set_token_tensor(toks, ntok)
    for t in range(W):
        dec.decode_step(pool[t % POOL])
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for t in range(K):
        dec.decode_step(pool[(W + t) % POOL])
    d1.record()
    barrier()
    dms = d0.elapsed_time(d1)
    # round-trip property on what was coded so far: decoded prefix == message prefix
    got = dec.out_bits.cpu().numpy().view(np.uint32)
    olen = dec.out_len.cpu().numpy()
    cur = st.cursor.cpu().numpy()
    rt_ok = True
    for r in range(0, B, max(1, B // 64)):
        n = int(min(cur[r], olen[r]))
        full, rem = n // 32, n % 32
        rt_ok &= bool(np.array_equal(got[r, :full], words[r, :full]))
        if rem:
            rt_ok &= bool((got[r, full] >> (32 - rem)) == (words[r, full] >> (32 - rem)))

    # end-to-end through the C ABI with HOST buffers: pinned logits -> H2D -> step -> tokens D2H
    e2e_steps = max(2, min(K, args.e2e_steps))
    host_logits = torch.empty((B, V), dtype=torch.float32, pin_memory=True)
    host_logits.copy_(pool[0])
    host_tok = torch.empty((B,), dtype=torch.int32, pin_memory=True)
    # two device buffers: the H2D copy of step t+1 (copy stream) runs under the coder step of step t
    dev_logits = [torch.empty((B, V), dtype=torch.float32, device=dev) for _ in range(2)]
    e2e = ArithmeticStreams(B, V, precision=PRECISION, temp=TEMP, topk=V, token_cap=e2e_steps + 4, device=dev)
    e2e.set_packed_messages(torch.from_numpy(words.view(np.int32)), torch.from_numpy(lens))
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    for ev in consumed:
        ev.record(main_stream)

    def e2e_step(t):
        b = t & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])              # the step that read this buffer two steps ago is done
            dev_logits[b].copy_(host_logits, non_blocking=True)
            copied[b].record(copy_stream)
        main_stream.wait_event(copied[b])
        e2e.encode_step(dev_logits[b])
        consumed[b].record(main_stream)
        host_tok.copy_(e2e.tokens[:, t], non_blocking=True)

    e2e_step(0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(1, 1 + e2e_steps):
        e2e_step(t)
    e1.record()
    barrier()
    ems = e0.elapsed_time(e1)

    # config 5: the comparison codecs at the same shape (rank 0; short legs, reported beside the headline)
    codecs = None
    if rank == 0 and not args.no_codecs:
        from neuralsteganography_b200.codecs import CodecStreams
        codecs = {}
        cw, ck = 2, max(2, min(K, args.codec_steps))
        for name, kind, kw in (("huffman_b3", "huffman", dict(param=3)), ("bins_b3", "bins", dict(param=3)),
                               ("rank", "rank", dict())):
            cs = CodecStreams(kind, B, V, token_cap=cw + ck + 2, device=dev, **kw)
            cs.set_packed_messages(torch.from_numpy(words.view(np.int32)), torch.from_numpy(lens))
            for t in range(cw):
                cs.encode_step(pool[t % POOL])
            torch.cuda.synchronize()
            c0 = int(cs.cursor.sum().item())
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for t in range(ck):
                cs.encode_step(pool[(cw + t) % POOL])
            k1.record()
            torch.cuda.synchronize()
            kms = k0.elapsed_time(k1)
            ctps = B * ck / (kms * 1e-3)
            codecs[name] = {"tokens_per_sec": ctps, "bits_per_token": (int(cs.cursor.sum().item()) - c0) / (B * ck),
                            "roofline_frac": ctps * 4 * V / 1e9 / peaks()[0], "steps": ck,
                            "flags": int((cs.status & 3).sum().item())}
            del cs

    # config 2's coder settings (temp 0.9, precision 26, topk 300: the rank form of the cutoff) at the same shape
    topk_leg = None
    if rank == 0 and not args.no_codecs:
        tk = ArithmeticStreams(B, V, precision=PRECISION, temp=0.9, topk=300, token_cap=K + W + 8, device=dev)
        tk.set_packed_messages(torch.from_numpy(words.view(np.int32)), torch.from_numpy(lens))
        tn = max(4, min(K, 4 * args.codec_steps))
        for t in range(W):
            tk.encode_step(pool[t % POOL])
        torch.cuda.synchronize()
        c0 = int(tk.cursor.sum().item())
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for t in range(tn):
            tk.encode_step(pool[(W + t) % POOL])
        k1.record()
        torch.cuda.synchronize()
        tms = k0.elapsed_time(k1)
        ttps = B * tn / (tms * 1e-3)
        # decode of the same cover tokens, and the round trip on a subset of the streams
        tdec = ArithmeticStreams(B, V, precision=PRECISION, temp=0.9, topk=300, token_cap=K + W + 8, device=dev)
        tdec.set_token_tensor(tk.tokens.clone(), tk.ntok.clone())
        for t in range(W):
            tdec.decode_step(pool[t % POOL])
        torch.cuda.synchronize()
        k0.record()
        for t in range(tn):
            tdec.decode_step(pool[(W + t) % POOL])
        k1.record()
        torch.cuda.synchronize()
        tdms = k0.elapsed_time(k1)
        tgot = tdec.out_bits.cpu().numpy().view(np.uint32)
        tolen = tdec.out_len.cpu().numpy()
        tcur = tk.cursor.cpu().numpy()
        trt = True
        for r in range(0, B, max(1, B // 64)):
            n = int(min(tcur[r], tolen[r]))
            full, rem = n // 32, n % 32
            trt &= bool(np.array_equal(tgot[r, :full], words[r, :full]))
            if rem:
                trt &= bool((tgot[r, full] >> (32 - rem)) == (words[r, full] >> (32 - rem)))
        topk_leg = {"workload": "same pool, temp 0.9, precision 26, topk 300", "tokens_per_sec": ttps,
                    "decode_tokens_per_sec": B * tn / (tdms * 1e-3), "roundtrip_ok": trt,
                    "bits_per_token": (int(tk.cursor.sum().item()) - c0) / (B * tn),
                    "roofline_frac": ttps * ALGO_BYTES_PER_TOKEN / 1e9 / peaks()[0], "steps": tn,
                    "kernel": "ac_topk_stream_kernel (csrc/ns_topk.cuh: the row is read once), then ac_fast_kernel on the rows it leaves",
                    "rows_left_to_row_resident_kernel": int(((tk.status & 16) != 0).sum().item()),
                    "rows_handed_to_exact_kernel": int(((tk.status & 4) != 0).sum().item())}
        del tk, tdec

    # final gather of the cover tokens (the only collective; outside the hot path)
    gather_ms = None
    if world > 1:
        from neuralsteganography_b200.sharding import gather_ragged
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gathered = gather_ragged(st.tokens, st.ntok, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        if rank == 0:
            assert len(gathered) == world * B
        t_all = torch.tensor([ms, dms, ems], device=dev, dtype=torch.float64)
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        ms, dms, ems = [float(x) for x in t_all.tolist()]
        cnt = torch.tensor([bits, live, int(rt_ok)], device=dev, dtype=torch.int64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        bits, live, rt_sum = [int(x) for x in cnt.tolist()]
        rt_ok = rt_sum == world

    # config 2's shape end to end through the provider API: host bits in -> host tokens out (GPT-2-small-shaped random-init
    # trunk in PyTorch + the coder at temp 0.9 / precision 26 / topk 300, whole loop under CUDA graphs)
    generation = None
    if rank == 0 and not args.no_generation:
        del pool
        torch.cuda.empty_cache()
        from neuralsteganography_b200.lm import B200ArithmeticLM, random_init_model
        gtok, gmodel = random_init_model("gpt2", seed=1234)
        glm = B200ArithmeticLM(gmodel.to(dev), gtok, device=dev, max_len=512)
        gB, gbits = args.gen_streams, 1024
        grng = np.random.Generator(np.random.PCG64(7))
        gmsgs = [grng.integers(0, 2, gbits).tolist() for _ in range(gB)]
        gq = {"temp": 0.9, "precision": 26, "topk": 300, "finish_sent": False}
        gctx = [50256, 464, 2068]
        gwarm = glm.encode_arithmetic_batch(gmsgs, gctx, quality=gq)         # warm-up: graph capture, cuBLAS plans
        glm.decode_arithmetic_batch(gwarm, gctx, quality=gq)                 # (both directions)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        covers = glm.encode_arithmetic_batch(gmsgs, gctx, quality=gq)
        torch.cuda.synchronize()
        t_enc = time.perf_counter() - t0
        t0 = time.perf_counter()
        back = glm.decode_arithmetic_batch(covers, gctx, quality=gq)
        torch.cuda.synchronize()
        t_dec = time.perf_counter() - t0
        ntok_g = sum(len(c) for c in covers)
        # the same loop with the trunk's GEMMs on the tensor cores (TF32; encoder and decoder share the setting)
        glm_t = B200ArithmeticLM(gmodel, gtok, device=dev, max_len=512, trunk_tf32=True)
        gw = glm_t.encode_arithmetic_batch(gmsgs, gctx, quality=gq)
        glm_t.decode_arithmetic_batch(gw, gctx, quality=gq)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        covers_t = glm_t.encode_arithmetic_batch(gmsgs, gctx, quality=gq)
        torch.cuda.synchronize()
        t_enc_t = time.perf_counter() - t0
        t0 = time.perf_counter()
        back_t = glm_t.decode_arithmetic_batch(covers_t, gctx, quality=gq)
        torch.cuda.synchronize()
        t_dec_t = time.perf_counter() - t0
        ntok_t = sum(len(c) for c in covers_t)
        tf32_leg = {"encode_tokens_per_sec": ntok_t / t_enc_t, "decode_tokens_per_sec": ntok_t / t_dec_t, "cover_tokens": ntok_t,
                    "roundtrip_ok": all(b[:gbits] == m for b, m in zip(back_t, gmsgs))}
        del glm_t
        generation = {"workload": "configs[1] shape: GPT-2-small random-init trunk (fp32), %d streams x %d message bits, temp 0.9, "
                                  "precision 26, topk 300; B200ArithmeticLM.encode/decode_arithmetic_batch, host lists in and out" % (gB, gbits),
                      "encode_tokens_per_sec": ntok_g / t_enc, "decode_tokens_per_sec": ntok_g / t_dec,
                      "encode_message_bits_per_sec": gB * gbits / t_enc, "cover_tokens": ntok_g,
                      "roundtrip_ok": all(b[:gbits] == m for b, m in zip(back, gmsgs)),
                      "trunk_tf32": tf32_leg}
        del glm, gmodel
        torch.cuda.empty_cache()

    # config 4: gpt2-fa-shaped random-init trunk (42001 tokens), 1024 streams, cover tokens -> messages: the step-wise
    # decode loop (one trunk step + one coder step per token, CUDA graphs) against the teacher-forced tiled decode
    config4 = None
    if rank == 0 and not args.no_generation:
        from neuralsteganography_b200.generation import StegoGenerator
        from neuralsteganography_b200.lm import random_init_model
        _t4, m4 = random_init_model("gpt2-fa")
        g4 = StegoGenerator(m4.to(dev), args.config4_streams, max_len=256, precision=26, temp=0.9, topk=300, use_graph=True, device=dev)
        c4 = torch.tensor([5, 11, 22])
        r4 = np.random.Generator(np.random.PCG64(44))
        nb4 = 512
        m4s = [r4.integers(0, 2, nb4).tolist() for _ in range(args.config4_streams)]
        t4 = g4.encode(c4, m4s, poll_every=16)
        n4 = sum(len(t) for t in t4)

        def timed(fn, reps=2):
            best, out = 1e9, None
            for _ in range(reps):
                torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            return best, out

        ts, seq = timed(lambda: g4.decode(c4, t4, poll_every=64))
        tt, til = timed(lambda: g4.decode_prefill(c4, t4, tile=32))
        config4 = {"workload": "configs[3]: gpt2-fa-shaped random-init trunk (V=42001, fp32), %d streams x %d message bits, %d cover "
                               "tokens, temp 0.9 / precision 26 / topk 300; host token lists in -> host bit lists out" % (args.config4_streams, nb4, n4),
                   "stepwise_decode_tokens_per_sec": n4 / ts, "tiled_decode_tokens_per_sec": n4 / tt, "tiled_over_stepwise": ts / tt,
                   "tile": 32, "all_messages_recovered": all(a[:nb4] == m and b[:nb4] == m for a, b, m in zip(seq, til, m4s))}
        del g4, m4
        torch.cuda.empty_cache()

    if rank == 0:
        tokens = world * B * K
        tps = tokens / (ms * 1e-3)
        peak, peak_src = peaks()
        kernel_s = ms * 1e-3 / K
        achieved = ALGO_BYTES_PER_TOKEN * B / kernel_s / 1e9
        line = {
            "metric": METRIC, "value": tps, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(B),
            "message_bits_per_sec": bits / (ms * 1e-3),
            "bits_per_token": bits / tokens,
            "decode_tokens_per_sec": tokens / (dms * 1e-3),
            "live_streams_at_end": live, "roundtrip_ok": bool(rt_ok),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_TOKEN * B, "traffic_source": NCU_TRAFFIC_SOURCE,
                         "peak_source": peak_src, "kernel": "ac_lean_kernel<unit_temp, ENC> (csrc/ns_lean.cuh)",
                         "algorithmic_bytes_per_token": ALGO_BYTES_PER_TOKEN},
            "cpu_baseline": cpu,
            "e2e": {"value": world * B * e2e_steps / (ems * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": B * V * 4, "d2h_bytes_per_step": B * 4,
                    "note": "host logits (pinned) -> H2D -> ns_ac_encode_step -> tokens D2H, per rank; every step copies its own logits, the copy of step t+1 overlaps the coder step of step t (two device buffers)"},
            "gpu_launches": 2 * K,   # per step: ac_lean_kernel + ac_step_kernel draining the hand-over queue
            "gather_ms": gather_ms,
            "codecs": codecs,
            "topk300": topk_leg,
            "generation": generation,
            "config4": config4,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS)
    ap.add_argument("--cpu-steps", type=int, default=400)
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--codec-steps", type=int, default=8)
    ap.add_argument("--no-codecs", action="store_true")
    ap.add_argument("--no-generation", action="store_true")
    ap.add_argument("--gen-streams", type=int, default=256)
    ap.add_argument("--config4-streams", type=int, default=1024)
    ap.add_argument("--strong", action="store_true", help="strong scaling: --streams is the total over all ranks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
