#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "mid_kernel_equals" > gpurun_out/t_ws2.log 2>&1; echo "ws all rc=$?"; tail -3 gpurun_out/t_ws2.log
C="python bench.py --config c3 --genes 4800 --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 300 $C > gpurun_out/c3_4800_ws.json 2> gpurun_out/c3_4800_ws.err; echo "c3 ws 32x6 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_ws.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
DEGNORM_B200_LIB=$PWD/degnorm_b200/libdegnorm_b200.c64.so timeout 300 $C > gpurun_out/c3_4800_ws64.json 2> gpurun_out/c3_4800_ws64.err; echo "c3 ws 64x3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_ws64.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
