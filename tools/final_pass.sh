#!/bin/bash
# Final measurement pass of a round on one B200 (run under gpurun from the repo root): bash tools/final_pass.sh r02
tag=${1:-rXX}; out=gpurun_out; mkdir -p $out
timeout 900 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
for cfg in cfg2 cfg3 cfg5; do timeout 900 python bench.py --config $cfg > $out/${tag}_bench_$cfg.json 2> $out/${tag}_bench_$cfg.err; done
KPROF=1 timeout 300 python tools/stage_times.py 32 > $out/${tag}_stage_times.txt 2>&1
timeout 200 python tools/aff_times.py 128 32 > $out/${tag}_aff_times.txt 2>&1
timeout 200 python tools/pn_times.py 128 512 32 > $out/${tag}_pn_times.txt 2>&1
# launch list of one bench step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 --pairs 16 --no-cpu > $out/${tag}_ncu_bench.log 2>&1
# ncu --set full of the trunk's tensor-core launches (first forward: layer 0 contraction + layers 1..12), 16 pairs
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tma" -c 13 -o $out/${tag}_conv \
  python tools/stage_times.py 16 > $out/${tag}_ncu_conv.log 2>&1
ls -la $out | grep $tag | tail -20
