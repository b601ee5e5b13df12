#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 2 0 1; do
  export DEGNORM_B200_LIB=$PWD/degnorm_b200/libdegnorm_b200.smn$v.so
  echo "=== variant $v"
  timeout 60 python tools/compare_kernels.py 48 8 1 12 2>&1 | tail -4 | cut -c1-300; echo "rc=$?"
  timeout 60 python tools/compare_kernels.py 30 16 1 12 2>&1 | tail -4 | cut -c1-300
  timeout 100 python bench.py --config c3 --genes 300 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_300_v$v.json 2> gpurun_out/c3_300_v$v.err; echo "c3 300 v$v rc=$?"; tail -3 gpurun_out/c3_300_v$v.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/c3_300_v$v.json')); print(d['value'], d['roofline']['frac'])" 2>/dev/null
done
unset DEGNORM_B200_LIB
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "wide_kernel_equals" > gpurun_out/t_wide2.log 2>&1; echo "wide all rc=$?"; tail -3 gpurun_out/t_wide2.log
timeout 300 python bench.py --config c5 --genes 296 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c5_296b.json 2> gpurun_out/c5_296b.err; echo "c5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c5_296b.json')); print(d['value'], d['roofline']['frac'], [b['end_ms'][0] for b in d['roofline']['buckets']])"
