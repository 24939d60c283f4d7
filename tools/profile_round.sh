#!/bin/bash
# One-call measurement pass of a round (run under gpurun from the repo root):  bash tools/profile_round.sh r02
# Leaves everything under gpurun_out/<tag>_*; the summaries worth judging are copied into profiles/ by hand.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
# 1. per-config bench lines (cfg4 is the default `python bench.py`)
for cfg in cfg2 cfg3; do
  timeout 900 python bench.py --config $cfg > $out/${tag}_bench_$cfg.json 2> $out/${tag}_bench_$cfg.err
done
timeout 900 python bench.py --config cfg5 > $out/${tag}_bench_cfg5.json 2> $out/${tag}_bench_cfg5.err
# 2. launch list of one bench step (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 --pairs 16 --no-cpu > $out/${tag}_ncu_bench.log 2>&1
# 3. ncu --set full: the three generated-operand contractions of the affinity stage, and the PointNet contractions
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_gen -s 3 -c 3 -o $out/${tag}_aff_gen \
  python tools/aff_times.py 128 8 > $out/${tag}_ncu_aff.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"newend_mean|link_logit" -s 2 -c 2 -o $out/${tag}_aff_hbm \
  python tools/aff_times.py 128 8 >> $out/${tag}_ncu_aff.log 2>&1
ls -la $out | tail -20
