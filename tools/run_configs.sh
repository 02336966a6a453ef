#!/bin/bash
# Runs the BASELINE.json configs beyond the headline one (config 2 = default bench.py) on one B200 and
# collects one JSON line per run in gpurun_out/configs.jsonl.  Usage: bash tools/run_configs.sh
out=gpurun_out/configs.jsonl; : > $out
run() { echo "# $*" >&2; timeout 900 python bench.py --no-cpu-baseline "$@" >> $out 2>> gpurun_out/configs.err || echo "{\"failed\": \"$*\"}" >> $out; }
# config 5: maxDist sweep on the human-scale genome, w=8 (default layout: triple sub-buckets up to maxDist 6)
run --slice-width 8 --max-dist 2 --steps 3 --warmup 1
run --slice-width 8 --max-dist 3 --steps 3 --warmup 1
run --slice-width 8 --max-dist 5 --steps 3 --warmup 1
run --slice-width 8 --max-dist 6 --guides 20000 --steps 2 --warmup 1
# config 5: slice-width sweep (w=10 takes the list scan over inline 32-bit residuals; w=4 the sub-bucket scan up to maxDist 4)
run --slice-width 10 --max-dist 3 --steps 3 --warmup 1
run --slice-width 10 --max-dist 4 --steps 3 --warmup 1
run --slice-width 4 --max-dist 4 --steps 3 --warmup 1
run --slice-width 4 --max-dist 4 --layout sig64 --guides 10000 --steps 2 --warmup 1
# config 4: repeat-rich mouse-scale genome (2.7 Gbp -> 506.25 M uniform sites + 2000 planted families of 5000 copies),
# half of the guides from the families, thresholds 0 (full scan) and 75 (early exit)
run --sites 506250000 --families 2000 --family-size 5000 --family-guides 0.5 --threshold 0 --steps 3 --warmup 1
run --sites 506250000 --families 2000 --family-size 5000 --family-guides 0.5 --threshold 75 --steps 3 --warmup 1
# the list-scan layouts on the headline workload: inline residuals with guide groups (bit-sliced blocks of 32), the
# north-star's ids+gather layout and the 64-bit inline layout; and the pure HBM stream (one guide per scan item)
run --layout res32 --steps 3 --warmup 1
run --layout res32 --max-group 1 --guides 20000 --steps 2 --warmup 1
run --layout gather --guides 20000 --steps 2 --warmup 1
run --layout sig64 --guides 20000 --steps 2 --warmup 1
# config 3 flavour: 1 M guides in one call, method and
run --guides 1000000 --steps 2 --warmup 1
# config 3: 10 M guides, method and (ten internal batches of 2^20 guides)
run --guides 10000000 --steps 1 --warmup 1
