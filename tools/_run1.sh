set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -k mid_kernel -x -q --timeout 300 > gpurun_out/pytest_mid.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_mid.log
B="python bench.py --config c3 --genes 2000 --steps 1 --warmup 1 --no-cpu --no-e2e"
for v in "w8:" "w4:--mid-warps 4" "w4h:--mid-warps 4 --mid-clusters 1:8192,2:16384,4:32768,8:65536,16:2000000000"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 600 $B $flags > gpurun_out/c3_$name.json 2> gpurun_out/c3_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/c3_$name.json')); print('$name', d['value'], d['roofline']['frac'], d['ms_per_step'])
except Exception as e: print('$name failed', e)
PY
done
