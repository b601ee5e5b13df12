#!/bin/bash
# C4 on a larger sample (800 of 5,000 genes: the 200-gene sample is bounded by its one 2 Mb gene on a 16-CTA cluster),
# then the single-matrix / fixture parity tests on the final host code
set -u
mkdir -p gpurun_out
timeout 200 python bench.py --config c4 --genes 800 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c4_800.json 2> gpurun_out/c4_800.err; echo "c4 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/c4_800.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], d['ms_per_step'], r['frac'], [(b['cluster'], b['resident'], b['genes'], b['max_cols'], b['end_ms'][1]) for b in r['buckets']])"
timeout 70 python -m pytest tests/test_gpu_parity.py -q -k "single_matrix or single_gene or golden_reference" > gpurun_out/final_targeted.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_targeted.log
