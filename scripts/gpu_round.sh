#!/bin/bash
# One gpurun call of a round: GPU tests, the bench line, phase clocks and the ncu evidence of the headline kernel.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r2a [tests] [bench] [clocks] [launches] [ncu]'
# Everything lands in gpurun_out/<tag>_*; copy what is to be judged into profiles/.
tag=${1:-r2}; shift
what=${*:-tests bench clocks launches ncu}
out=gpurun_out
mkdir -p $out
small="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-generation --streams 592"
for w in $what; do
  case $w in
    tests)    timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gputest.log 2>&1; tail -5 $out/${tag}_gputest.log ;;
    smoke)    timeout 300 python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log ;;
    bench)    timeout 900 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; tail -c 3000 $out/${tag}_bench_1gpu.json ;;
    quick)    timeout 600 python bench.py --steps 50 --no-cpu-baseline --no-generation --no-codecs > $out/${tag}_quick.json 2> $out/${tag}_quick.err; tail -c 1500 $out/${tag}_quick.json; tail -3 $out/${tag}_quick.err ;;
    clocks)   DUO=0 timeout 300 python scripts/phase_clocks.py > $out/${tag}_phase_clocks.txt 2>&1; cat $out/${tag}_phase_clocks.txt ;;
    launches) $small > $out/${tag}_plain.log 2>&1 &&
              timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv $small > $out/${tag}_ncu_launches.log 2>&1
              tail -2 $out/${tag}_ncu_launches.log ;;
    ncu)      $small --no-codecs > $out/${tag}_plain2.log 2>&1 &&
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-ac_lean_kernel} -s 4 -c 1 -f -o $out/${tag}_prof $small --no-codecs > $out/${tag}_ncu_full.log 2>&1
              tail -2 $out/${tag}_ncu_full.log ;;
    ncu_rank) $small > $out/${tag}_plain3.log 2>&1 &&
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:codec_stream_kernel -s 6 -c 1 -f -o $out/${tag}_rank_prof $small > $out/${tag}_ncu_rank.log 2>&1
              tail -2 $out/${tag}_ncu_rank.log ;;
    soak)     STREAMS=${SOAK_STREAMS:-1024} STEPS=${SOAK_STEPS:-40} timeout 900 python scripts/soak_fast_vs_exact.py > $out/${tag}_soak.txt 2>&1; tail -12 $out/${tag}_soak.txt ;;
  esac
done
