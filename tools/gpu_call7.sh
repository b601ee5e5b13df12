#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 60 python tools/compare_kernels.py 48 8 1 12 2>&1 | tail -4 | cut -c1-200
timeout 60 python tools/compare_kernels.py 30 16 1 12 2>&1 | tail -4 | cut -c1-200
timeout 100 python bench.py --config c3 --genes 300 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_300.json 2> gpurun_out/c3_300.err; echo "c3 300 rc=$?"; tail -3 gpurun_out/c3_300.err | cut -c1-300
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "mid_kernel_equals" > gpurun_out/t_ws2.log 2>&1; echo "ws all rc=$?"; tail -3 gpurun_out/t_ws2.log
timeout 300 python bench.py --config c3 --genes 4800 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_4800_ws.json 2> gpurun_out/c3_4800_ws.err; echo "c3 ws rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_ws.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
W="python bench.py --config c5 --genes 148 --max-len 2500 --steps 1 --warmup 0 --no-cpu --no-e2e"
timeout 200 $W > gpurun_out/plain_wide.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmfoa_wide -s 2 -c 1 -o gpurun_out/prof_wide -f $W > gpurun_out/ncu_wide.log 2>&1; echo "ncu wide rc=$?"; tail -2 gpurun_out/ncu_wide.log
