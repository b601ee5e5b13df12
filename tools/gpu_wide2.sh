#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "wide_kernel_equals or seed_p200" > gpurun_out/t_wide.log 2>&1; echo "wide tests rc=$?"; tail -2 gpurun_out/t_wide.log
timeout 300 python bench.py --config c5 --genes 296 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c5_296d.json 2> gpurun_out/c5_296d.err; echo "c5 296 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c5_296d.json')); print(d['value'], d['roofline']['frac'], [b['end_ms'][0] for b in d['roofline']['buckets']])"
timeout 600 python bench.py --config c5 --genes 1184 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c5_1184.json 2> gpurun_out/c5_1184.err; echo "c5 1184 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c5_1184.json')); print(d['value'], d['roofline']['frac'], d['roofline']['hbm']['frac'], [b['end_ms'][0] for b in d['roofline']['buckets']])"
