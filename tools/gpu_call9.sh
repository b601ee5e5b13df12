#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "mid_kernel_equals" > gpurun_out/t_ws2.log 2>&1; echo "ws all rc=$?"; tail -3 gpurun_out/t_ws2.log
C="python bench.py --config c3 --genes 4800 --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 300 $C > gpurun_out/c3_4800_ws.json 2> gpurun_out/c3_4800_ws.err; echo "c3 ws rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_ws.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
N="python bench.py --config c3 --genes 592 --max-len 4000 --steps 1 --warmup 0 --no-cpu --no-e2e"
timeout 200 $N > gpurun_out/plain_mid.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmfoa_mid -s 2 -c 1 -o gpurun_out/prof_mid_ws2 -f $N > gpurun_out/ncu_mid.log 2>&1; echo "ncu mid rc=$?"; tail -2 gpurun_out/ncu_mid.log
