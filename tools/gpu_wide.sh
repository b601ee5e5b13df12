#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "wide_kernel_equals or seed_p100 or seed_p200 or (sample_counts and (70 or 130 or 200 or 256))" > gpurun_out/t_wide.log 2>&1; echo "wide tests rc=$?"; tail -3 gpurun_out/t_wide.log
timeout 300 python bench.py --config c5 --genes 296 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c5_296c.json 2> gpurun_out/c5_296c.err; echo "c5 296 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c5_296c.json')); print(d['value'], d['roofline']['frac'], [b['end_ms'][0] for b in d['roofline']['buckets']])"
timeout 200 python bench.py --config c5 --genes 74 --max-len 1200 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c5_mini74.json 2> gpurun_out/c5_mini74.err; echo "c5 mini rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c5_mini74.json')); print(d['value'], d['roofline']['frac'])"
