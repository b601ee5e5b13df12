set -u
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config c3 --genes 2000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c3_2000.json 2> gpurun_out/c3_2000.err; echo "c3 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/c3_2000.json')); print('c3', d['value'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['buckets'])
PY
B="python bench.py --config c3 --genes 592 --max-len 4000 --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 600 $B > gpurun_out/c3_592.json 2> gpurun_out/c3_592.err; echo "plain rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c3_592.json')); print(d['value'], d['roofline']['frac'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nmfoa_mid -s 6 -c 1 -o gpurun_out/prof_mid_w8_tstore -f $B > gpurun_out/ncu_mid_w8_ts.log 2>&1; echo "ncu rc=$?"
