#!/bin/bash
# Round-2 evidence pass on one B200: parity tests, the headline bench line (C3 sample), launch list + DRAM traffic of
# one step of the same workload, ncu --set full of the mid-p kernel.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_c3.json
ONE="python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e"
timeout 300 $ONE > gpurun_out/plain_one.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"nmfoa|sums_|_apply|estimates" -c 60 --csv --log-file gpurun_out/launches_c3.csv $ONE > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
python tools/traffic_from_ncu.py gpurun_out/launches_c3.csv c3 4800 5 nmfoa_mid gpurun_out/traffic_c3.json
N="python bench.py --config c3 --genes 592 --max-len 4000 --steps 1 --warmup 0 --no-cpu --no-e2e"
timeout 200 $N > gpurun_out/plain_mid.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmfoa_mid -s 2 -c 1 -o gpurun_out/prof_mid_final -f $N > gpurun_out/ncu_mid.log 2>&1; echo "ncu mid rc=$?"; tail -2 gpurun_out/ncu_mid.log
ls -la gpurun_out | tail -12
