#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/probe_peaks.py > gpurun_out/probes.json 2> gpurun_out/probes.err; echo "probes rc=$?"; cat gpurun_out/probes.json
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "mid_kernel or production or sample_counts" > gpurun_out/pytest_mid.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_mid.log
timeout 600 python bench.py --config c3 --genes 4800 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_4800_ws.json 2> gpurun_out/c3_4800_ws.err; echo "c3 ws rc=$?"; cut -c1-200 gpurun_out/c3_4800_ws.json; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_ws.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
timeout 600 python bench.py --config c3 --genes 4800 --steps 1 --warmup 1 --no-cpu --no-e2e --mid-warps 8 > gpurun_out/c3_4800_w8.json 2> gpurun_out/c3_4800_w8.err; echo "c3 w8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_4800_w8.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
