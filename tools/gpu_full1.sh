#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 2000 > gpurun_out/clocks_full1.csv &
SMI=$!
timeout 1500 python bench.py --config c3 --genes 60000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_full_1gpu.json 2> gpurun_out/c3_full_1gpu.err; echo "c3 full 1 gpu rc=$?"
kill $SMI
tail -3 gpurun_out/c3_full_1gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/c3_full_1gpu.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'], d['clocks'])"
