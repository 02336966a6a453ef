#!/bin/bash
# BASELINE.json configs[3] as SURVEY.md 8d (C4) states it: 2.7 Gbp-scale index (506.25 M uniform sites) + 2 000 near-repeat
# families of 100-100 000 copies (log-uniform), per-base substitution rate 0-15 %, + 1 % low-complexity tract sites; half of the
# guides from family members; thresholds 0 (full scan) and 75 (early exit).  Each line carries the list-length histogram,
# hits/guide, the early-exit fraction, the CPU baseline (unmodified reference) and the line parity against it.
out=${1:-gpurun_out/config4.jsonl}; : > $out
for thr in 0 75; do
  timeout 1200 python bench.py --sites 506250000 --families 2000 --family-size 100 --family-size-max 100000 --low-complexity 0.01 \
      --family-guides 0.5 --guides 100000 --threshold $thr --steps 3 --warmup 3 --cpu-guides 400 >> $out 2>> gpurun_out/config4.err \
      || echo "{\"failed\": \"threshold $thr\"}" >> $out
done
