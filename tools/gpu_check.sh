#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "mid_kernel_equals or seed_p17 or seed_p48 or (sample_counts and (17 or 33 or 48))" > gpurun_out/t_mid.log 2>&1; echo "mid tests rc=$?"; tail -3 gpurun_out/t_mid.log
C="python bench.py --config c3 --genes 4800 --steps 2 --warmup 1 --no-cpu --no-e2e"
timeout 300 $C > gpurun_out/c3_4800_w8.json 2> gpurun_out/c3_4800_w8.err; echo "c3 w8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_w8.json')); print(d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])"
