#!/bin/bash
# C4 (long genes, p = 12, streamed cluster tier of the small-p kernel): bench sample + ncu of the shipped kernel
set -u
mkdir -p gpurun_out
timeout 400 python bench.py --config c4 --genes 200 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c4_200.json 2> gpurun_out/c4_200.err; echo "c4 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c4_200.json')); r=d['roofline']; print(d['value'], r['frac'], r['bound'][:40], [(b['cluster'], b['resident'], b['genes'], b['end_ms'][0]) for b in r['buckets']])"
N="python bench.py --config c4 --genes 60 --steps 1 --warmup 0 --no-cpu --no-e2e"
timeout 300 $N > gpurun_out/plain_c4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:nmfoa_small -s 4 -c 1 -o gpurun_out/prof_small_streamed -f $N > gpurun_out/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"; tail -2 gpurun_out/ncu_c4.log
