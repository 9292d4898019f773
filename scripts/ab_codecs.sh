for v in "" ${AB_VARIANTS:-}; do
  if [ -n "$v" ]; then export NS_CODER_LIB=gpurun_bin/libns_$v.so; else unset NS_CODER_LIB; fi
  python bench.py --steps 10 --no-cpu-baseline --no-generation 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['codecs']
print('${v:-base}', 'rank %.2f huff %.2f bins %.2f topk300 %.2f' % (c['rank']['tokens_per_sec']/1e6, c['huffman_b3']['tokens_per_sec']/1e6, c['bins_b3']['tokens_per_sec']/1e6, d['topk300']['tokens_per_sec']/1e6))"
done
