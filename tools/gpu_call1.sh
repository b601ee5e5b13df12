#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/probe_peaks.py > gpurun_out/probes.json 2> gpurun_out/probes.err; echo "probes rc=$?"; cat gpurun_out/probes.json
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config c3 --genes 4800 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_4800_base.json 2> gpurun_out/c3_4800_base.err; echo "c3 rc=$?"; cut -c1-600 gpurun_out/c3_4800_base.json
