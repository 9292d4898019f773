"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/ref_harness.py).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Every case records the inputs' recipe (numpy PCG64 seeds + a CRC32 of the logits
pool so RNG drift is detected), the reference's outputs (tokens from the live
``encode_arithmetic``, bits from the live ``decode_arithmetic``) and the oracle's
per-step trace.  The script asserts oracle == reference on every case before
writing, which is what pins ``oracle/ac_oracle.py``.
"""

from __future__ import annotations

import json
import os
import sys
import zlib

import numpy as np

from . import ac_oracle as O
from . import ref_harness as H
from . import codecs_oracle as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


from .inputs import logits_pool, make_distinct, message_bits, rows_for  # noqa: E402,F401


AC_CASES = [
    # name, V, T, scale, precision, topk, temp, streams, bits
    dict(name="ac_v2048_p26_k300_t09", V=2048, T=32, scale=3.0, precision=26, topk=300, temp=0.9, streams=6, bits=192),
    dict(name="ac_v2048_p16_full_t10", V=2048, T=32, scale=3.0, precision=16, topk=50000, temp=1.0, streams=6, bits=160),
    dict(name="ac_v2048_p32_full_t13", V=2048, T=32, scale=2.0, precision=32, topk=2048, temp=1.3, streams=4, bits=200),
    dict(name="ac_v2048_p40_k60000", V=2048, T=32, scale=3.0, precision=40, topk=60000, temp=1.0, streams=3, bits=200),
    dict(name="ac_v2048_p8_k5_t07", V=2048, T=32, scale=1.0, precision=8, topk=5, temp=0.7, streams=4, bits=96),
    dict(name="ac_v50257_p26_k300_t09", V=50257, T=12, scale=3.0, precision=26, topk=300, temp=0.9, streams=3, bits=160),
    dict(name="ac_v50257_p26_full_t10", V=50257, T=12, scale=3.0, precision=26, topk=50257, temp=1.0, streams=4, bits=256),
    dict(name="ac_v42001_p16_k50000", V=42001, T=8, scale=2.5, precision=16, topk=50000, temp=1.0, streams=2, bits=128),
]


def run_ac_case(cfg, seed_base: int):
    pool = logits_pool(seed_base, cfg["T"], cfg["V"], cfg["scale"])
    crc = zlib.crc32(pool.tobytes()) & 0xFFFFFFFF
    kw = dict(temp=cfg["temp"], precision=cfg["precision"], topk=cfg["topk"])
    out = dict(pool_seed=seed_base, pool_crc=crc)
    msgs, toks, dec, traces, stats = [], [], [], [], []
    for s in range(cfg["streams"]):
        rows = rows_for(pool, s)
        msg = message_bits(seed_base * 1000 + s, cfg["bits"] - 8 * (s % 3))
        ref_tok, ref_stats = H.ref_encode_arithmetic(rows, msg.tolist(), **kw)
        ref_bits = H.ref_decode_arithmetic(rows, ref_tok, **kw)
        res = O.encode_stream(rows, msg.tolist(), want_stats=True, **kw)
        obits, otrace = O.decode_stream(rows, ref_tok, keep_trace=True, **kw)
        assert res.tokens == ref_tok, (cfg["name"], s, "encode tokens differ from reference")
        assert obits == ref_bits, (cfg["name"], s, "decode bits differ from reference")
        assert ref_bits[: len(msg)] == msg.tolist(), (cfg["name"], s, "reference round trip failed")
        np.testing.assert_allclose([res.avg_nll, res.avg_kl, res.words_per_bit, res.avg_hq], ref_stats, rtol=1e-9)
        msgs.append(msg)
        toks.append(np.asarray(ref_tok, dtype=np.int32))
        dec.append(np.asarray(ref_bits, dtype=np.uint8))
        traces.append(np.asarray([[t.new_bottom, t.new_top, t.nbits, t.lo, t.hi, t.k, t.selection] for t in res.trace],
                                 dtype=np.int64))
        stats.append(np.asarray(ref_stats, dtype=np.float64))
    for s in range(cfg["streams"]):
        out["msg_%d" % s] = msgs[s]
        out["tokens_%d" % s] = toks[s]
        out["decoded_%d" % s] = dec[s]
        out["trace_%d" % s] = traces[s]
        out["stats_%d" % s] = stats[s]
    return out


def main() -> None:
    if not H.available():
        sys.exit("reference not mounted; golden vectors can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    meta = {"ac": []}
    for idx, cfg in enumerate(AC_CASES):
        data = run_ac_case(cfg, 1000 + idx)
        np.savez_compressed(os.path.join(OUT, cfg["name"] + ".npz"), **data)
        meta["ac"].append(cfg)
        print("ac", cfg["name"], "ok; tokens/stream", [len(data["tokens_%d" % s]) for s in range(cfg["streams"])])
    meta.update(K.make_codec_goldens(OUT, logits_pool, message_bits, rows_for))
    with open(os.path.join(OUT, "cases.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
