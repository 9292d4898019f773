#!/bin/bash
# A/B of experimental library builds (scripts/build_variant.sh) on one GPU box:
#   gpurun -- 'bash scripts/ab_variants.sh base a ab ...'   -> one line per build: encode / decode M tok/s, round trip
# Correctness of a candidate: NS_CODER_LIB=... pytest + soak (second block, only for names after "--check").
check=0
for name in "$@"; do
  if [ "$name" == "--check" ]; then check=1; continue; fi
  export NS_CODER_LIB=$PWD/gpurun_bin/libns_$name.so
  if [ $check == 0 ] && [ -n "$AB_TOPK" ]; then
    printf "%-8s " $name; timeout 240 python scripts/bench_topk.py 2>&1 | head -1
  elif [ $check == 0 ]; then
    timeout 240 python bench.py --steps ${AB_STEPS:-40} --no-cpu-baseline --no-generation --no-codecs > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
    python - $name <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/ab_%s.json' % n).read().strip().splitlines()[-1])
    print("%-8s enc %.3f M tok/s (frac %.4f)  dec %.3f M  roundtrip %s" % (n, d['value'] / 1e6, d['roofline']['frac'], d['decode_tokens_per_sec'] / 1e6, d['roundtrip_ok']))
except Exception as e:
    print(n, "no bench line", e)
PY
  else
    echo "== check $name"
    timeout 600 python -m pytest tests/test_ac_gpu.py tests/test_configs_gpu.py -m gpu -x -q 2>&1 | tail -2
    STREAMS=${SOAK_STREAMS:-1024} STEPS=${SOAK_STEPS:-24} timeout 600 python scripts/soak_fast_vs_exact.py 2>&1 | tail -3
  fi
done
