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
# dram__bytes_read.sum + dram__bytes_write.sum of ac_fast_kernel from one `ncu --set full` capture
# (592 rows: 119.27 MB + 7.20 MB, profiles/r1_final_fast_ncu_summary.txt), per row
NCU_DRAM_BYTES_PER_TOKEN = (119269888 + 7197184) / 592
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
def _cpu_worker(args):
    seed, steps = args
    from oracle import ac_oracle as O
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = [rng.standard_normal(V, dtype=np.float32) * np.float32(3.0) for _ in range(4)]
    msg = rng.integers(0, 2, MSG_BITS).tolist()
    t0 = time.perf_counter()
    res = O.encode_stream(lambda t: rows[t % 4], msg, temp=TEMP, precision=PRECISION, topk=V,
                          max_steps=steps, keep_trace=False)
    return len(res.tokens), res.bits_consumed, time.perf_counter() - t0


def cpu_codec_throughput(steps_per_stream: int, procs: int):
    """tokens/s of the CPU codec with `procs` independent streams in parallel."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        out = pool.map(_cpu_worker, [(100 + i, steps_per_stream) for i in range(procs)])
    wall = time.perf_counter() - t0
    toks = sum(o[0] for o in out)
    bits = sum(o[1] for o in out)
    busy = max(o[2] for o in out)
    return toks / busy, bits / busy, toks, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per = max(64, 4 * args.steps)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_codec_throughput(4, cores)
    tps, bps, toks, wall = cpu_codec_throughput(per, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * STREAMS / tps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU codec = oracle port of code_base/arithmetic.py (the reference is "
                   "not present on the GPU box); ms_per_step = time this host needs for one 4096-stream step"},
        "message_bits_per_sec": bps,
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": "%d independent streams x %d encode steps, V=50257, one process per core" % (cores, per)},
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
        ctps, cbps, ctoks, cwall = cpu_codec_throughput(args.cpu_steps, cores)
        cpu = {"value": ctps, "unit": "tokens/s", "cores": cores, "kind": "port",
               "message_bits_per_sec": cbps,
               "sample": "%d independent streams x %d encode steps of the same workload (V=50257, precision 26, "
                         "full distribution), one process per core; %d tokens in %.1f s wall" % (cores, args.cpu_steps, ctoks, cwall)}
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
        topk_leg = {"workload": "same pool, temp 0.9, precision 26, topk 300", "tokens_per_sec": ttps,
                    "bits_per_token": (int(tk.cursor.sum().item()) - c0) / (B * tn),
                    "roofline_frac": ttps * ALGO_BYTES_PER_TOKEN / 1e9 / peaks()[0], "steps": tn,
                    "rows_handed_to_exact_kernel": int(((tk.status & 4) != 0).sum().item())}
        del tk

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
            "config": {"workload": WORKLOAD, "streams_per_gpu": B, "vocab": V, "precision": PRECISION, "temp": TEMP,
                       "topk": V, "l2": "inputs larger than L2: 4-entry logits pool, 823 MB per step",
                       "parallelism": "streams sharded over ranks, no collective in the loop"},
            "message_bits_per_sec": bits / (ms * 1e-3),
            "bits_per_token": bits / tokens,
            "decode_tokens_per_sec": tokens / (dms * 1e-3),
            "live_streams_at_end": live, "roundtrip_ok": bool(rt_ok),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_TOKEN * B, "traffic_source": "ncu capture of 592 rows, scaled per row",
                         "peak_source": peak_src, "kernel": "ac_fast_kernel<unit_temp, ENC>",
                         "algorithmic_bytes_per_token": ALGO_BYTES_PER_TOKEN},
            "cpu_baseline": cpu,
            "e2e": {"value": world * B * e2e_steps / (ems * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": B * V * 4, "d2h_bytes_per_step": B * 4,
                    "note": "host logits (pinned) -> H2D -> ns_ac_encode_step -> tokens D2H, per rank; every step copies its own logits, the copy of step t+1 overlaps the coder step of step t (two device buffers)"},
            "gpu_launches": 2 * K,   # per step: ac_fast_kernel + ac_step_kernel draining the hand-over queue
            "gather_ms": gather_ms,
            "codecs": codecs,
            "topk300": topk_leg,
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
    ap.add_argument("--strong", action="store_true", help="strong scaling: --streams is the total over all ranks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
