#!/bin/bash
# full-size C3 (60,000 genes x 48 samples) sharded over the 8 GPUs of one box, device-resident
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 2000 > gpurun_out/clocks_full8.csv &
SMI=$!
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $T bench.py --gpus 8 --genes 60000 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_full_8gpu.json 2> gpurun_out/c3_full_8gpu.err; echo "c3 full 8 gpu rc=$?"
tail -3 gpurun_out/c3_full_8gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/c3_full_8gpu.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], [ (r['genes'], round(r['bs_ms_per_step']), round(r['wait_ms_per_step'])) for r in d['ranks']])"
kill $SMI
