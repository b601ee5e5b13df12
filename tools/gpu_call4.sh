#!/bin/bash
# risky new kernels first, each under a short timeout
set -u
mkdir -p gpurun_out
T="timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120"
$T -k "mid_kernel_equals and 48-1-0" > gpurun_out/t_ws1.log 2>&1; echo "ws 48-1 rc=$?"; tail -3 gpurun_out/t_ws1.log
$T -k "mid_kernel_equals" > gpurun_out/t_ws2.log 2>&1; echo "ws all rc=$?"; tail -3 gpurun_out/t_ws2.log
$T -k "wide_kernel_equals and 200-1" > gpurun_out/t_wide1.log 2>&1; echo "wide 200-1 rc=$?"; tail -3 gpurun_out/t_wide1.log
$T -k "wide_kernel_equals" > gpurun_out/t_wide2.log 2>&1; echo "wide all rc=$?"; tail -3 gpurun_out/t_wide2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "production or sample_counts" > gpurun_out/t_prod.log 2>&1; echo "production rc=$?"; tail -8 gpurun_out/t_prod.log
timeout 300 python bench.py --config c3 --genes 4800 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_4800_ws.json 2> gpurun_out/c3_4800_ws.err; echo "c3 ws rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_ws.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
timeout 600 python bench.py --config c5 --genes 296 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c5_296.json 2> gpurun_out/c5_296.err; echo "c5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c5_296.json')); print(d['value'], d['roofline']['frac'], d['roofline'].get('fp64'), d['roofline']['phases_ms_per_step'])"
