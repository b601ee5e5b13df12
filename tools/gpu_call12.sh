#!/bin/bash
set -u
mkdir -p gpurun_out
export DEGNORM_B200_LIB=$PWD/degnorm_b200/libdegnorm_b200.pipe.so
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "mid_kernel_equals or seed_p48" > gpurun_out/t_pipe.log 2>&1; echo "pipe tests rc=$?"; tail -2 gpurun_out/t_pipe.log
C="python bench.py --config c3 --genes 4800 --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 300 $C > gpurun_out/c3_4800_pipe.json 2> gpurun_out/c3_4800_pipe.err; echo "c3 pipe rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_pipe.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
unset DEGNORM_B200_LIB
timeout 300 $C > gpurun_out/c3_4800_w8.json 2> gpurun_out/c3_4800_w8.err; echo "c3 w8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_w8.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
