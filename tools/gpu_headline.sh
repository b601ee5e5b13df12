#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); r=d['roofline']; print(d['value'], r['frac'], r['traffic'], d['e2e']['value'], {k: v['value'] for k, v in d['e2e']['variants'].items()}, d['cpu_baseline']['value'], d['clocks'])"
