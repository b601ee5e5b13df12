#!/bin/bash
# One GPU-box pass for a round: parity tests, the headline bench line, launch list, ncu captures.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
SMALL="python bench.py --genes 4000 --steps 1 --warmup 1 --no-cpu --no-e2e"
$SMALL > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nmfoa|sums_|_apply|estimates" -s 59 -c 59 --csv --log-file gpurun_out/launches_c2.csv $SMALL > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:nmfoa_small -s 47 -c 1 -o gpurun_out/prof_small_t420 -f $SMALL > gpurun_out/ncu_t420.log 2>&1; echo "ncu t420 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:nmfoa_small -s 51 -c 1 -o gpurun_out/prof_small_t96 -f $SMALL > gpurun_out/ncu_t96.log 2>&1; echo "ncu t96 rc=$?"
C3="python bench.py --config c3 --genes 400 --steps 1 --warmup 1 --no-cpu --no-e2e"
$C3 > gpurun_out/c3_400.json 2> gpurun_out/c3_400.err && \
ncu --set full --clock-control none --import-source on -k regex:nmfoa_kernel -s 12 -c 1 -o gpurun_out/prof_tiled_p48 -f $C3 > gpurun_out/ncu_p48.log 2>&1; echo "ncu p48 rc=$?"
timeout 900 python bench.py --config c3 --genes 2000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_2000.json 2> gpurun_out/c3_2000.err; echo "c3 rc=$?"
ls -la gpurun_out
