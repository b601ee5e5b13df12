#!/bin/bash
# One GPU-box pass for a round: parity tests, the headline bench line, launch list, traffic, ncu captures.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
# one step of the headline workload (20,000 genes), warm-up 1: 50 launches per step (init 3 + 5 x (9 buckets) + ...)
ONE="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
$ONE > gpurun_out/plain_one.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"nmfoa|sums_|_apply|estimates" -s 50 -c 50 --csv --log-file gpurun_out/launches_c2.csv $ONE > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
SMALL="python bench.py --genes 4000 --steps 1 --warmup 1 --no-cpu --no-e2e"
$SMALL > gpurun_out/plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nmfoa_small -s 47 -c 1 -o gpurun_out/prof_small_a -f $SMALL > gpurun_out/ncu_a.log 2>&1; echo "ncu a rc=$?"
ncu --set full --clock-control none --import-source on -k regex:nmfoa_small -s 51 -c 1 -o gpurun_out/prof_small_b -f $SMALL > gpurun_out/ncu_b.log 2>&1; echo "ncu b rc=$?"
timeout 900 python bench.py --config c3 --genes 2000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_2000.json 2> gpurun_out/c3_2000.err; echo "c3 rc=$?"
timeout 900 python bench.py --config c4 --genes 200 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c4_200.json 2> gpurun_out/c4_200.err; echo "c4 rc=$?"
timeout 900 python bench.py --config c1 --steps 2 --warmup 1 --no-cpu > gpurun_out/c1.json 2> gpurun_out/c1.err; echo "c1 rc=$?"
# (C5, p = 200, runs on the untuned tiled kernel: 60 genes take ~5.5 minutes per run; measured once, see profiles/)
ls -la gpurun_out
