#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/probe_peaks.py > gpurun_out/probes.json 2> gpurun_out/probes.err; cat gpurun_out/probes.json
for cfg in "72 16" "72 2" "72 8" "200 16" "56 16"; do
  timeout 90 python tools/compare_kernels.py $cfg 1 12 2>&1 | tail -4
done
for cfg in "30 16" "48 16" "48 8" "30 2"; do
  timeout 90 python tools/compare_kernels.py $cfg 1 12 2>&1 | tail -6
done
timeout 200 python bench.py --config c3 --genes 600 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_600_ws.json 2> gpurun_out/c3_600_ws.err; echo "c3 600 ws rc=$?"; tail -5 gpurun_out/c3_600_ws.err; python -c "
import json; d=json.load(open('gpurun_out/c3_600_ws.json')); print(d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])"
