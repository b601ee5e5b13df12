set -u
mkdir -p gpurun_out
B="python bench.py --config c3 --genes 2000 --steps 1 --warmup 1 --no-cpu --no-e2e"
DEGNORM_B200_LIB=$PWD/degnorm_b200/libdegnorm_b200.f3.so timeout 600 python -m pytest tests/test_gpu_parity.py -k mid_kernel -x -q --timeout 300 > gpurun_out/pytest_mid.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_mid.log
for v in f3 f3w4; do
  lib=degnorm_b200/libdegnorm_b200.f3.so; flags=""
  if [ $v = f3w4 ]; then flags="--mid-warps 4"; fi
  DEGNORM_B200_LIB=$PWD/$lib timeout 600 $B $flags > gpurun_out/c3_$v.json 2> gpurun_out/c3_$v.err; echo "$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/c3_$v.json')); print('$v', d['value'], d['roofline']['frac'], d['ms_per_step'])
except Exception as e: print('$v failed', e)
PY
done
