#!/bin/bash
# two GPUs: the tests that need two devices (NCCL twin of run_gene_nmfoa_mpi, non-current device) and the default
# bench at N = 2 (strong scaling, end-to-end leg included)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 500 -k "nccl or non_current" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"; tail -3 gpurun_out/pytest_2gpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $T bench.py --gpus 2 --steps 2 --warmup 1 --no-variants > gpurun_out/c3_sample_2gpu.json 2> gpurun_out/c3_sample_2gpu.err; echo "c3 sample 2 gpu rc=$?"; tail -3 gpurun_out/c3_sample_2gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/c3_sample_2gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'] if d['e2e'] else None, [ (r['genes'], round(r['bs_ms_per_step']), round(r['wait_ms_per_step'])) for r in d['ranks']])"
