#!/bin/bash
# A/B of the small-p kernel's Gram-round unrolling (bitwise-neutral tuning knobs SMALL_STREAM_UNROLL / SMALL_RES_UNROLL,
# variant libraries built by degnorm_b200/build.py with DEGNORM_B200_VARIANT): C4 sample per variant, the streamed /
# cluster / reference-fixture parity tests on the fastest one, then C2 with the resident-tier variant if time is left.
set -u
mkdir -p gpurun_out
D=degnorm_b200
run_c4() {   # $1 = tag, $2 = library
  DEGNORM_B200_LIB=$2 timeout 35 python bench.py --config c4 --genes 200 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/c4u_$1.json 2> gpurun_out/c4u_$1.err
  echo "c4 $1 rc=$? $(python -c "
import json; d=json.loads(open('gpurun_out/c4u_$1.json').read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['roofline']['frac'],4))" 2>&1 | tail -1) t=$SECONDS"
}
run_c4 u4 $PWD/$D/libdegnorm_b200.so
run_c4 u8 $PWD/$D/libdegnorm_b200.u8.so
run_c4 u16 $PWD/$D/libdegnorm_b200.u16.so
BEST=$(python - <<'PY'
import json
best, bv = "u4", 0.0
for t in ("u4", "u8", "u16"):
    try:
        v = json.loads(open("gpurun_out/c4u_%s.json" % t).read().strip().splitlines()[-1])["value"]
    except Exception:
        v = 0.0
    if v > bv * 1.01:
        best, bv = t, v
print(best)
PY
)
echo "best=$BEST"
LIB=$PWD/$D/libdegnorm_b200.so; [ "$BEST" != "u4" ] && LIB=$PWD/$D/libdegnorm_b200.$BEST.so
DEGNORM_B200_LIB=$LIB timeout 60 python -m pytest tests/test_gpu_parity.py -q -x -k "streamed or cluster_path or seed_p12_long or golden_reference or mixed_lengths" > gpurun_out/unroll_parity_$BEST.log 2>&1
echo "parity($BEST) rc=$? $(tail -1 gpurun_out/unroll_parity_$BEST.log) t=$SECONDS"
if [ $SECONDS -lt 120 ]; then
  for v in so r8.so; do
    DEGNORM_B200_LIB=$PWD/$D/libdegnorm_b200.$v timeout 32 python bench.py --config c2 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/c2u_$v.json 2> gpurun_out/c2u_$v.err
    echo "c2 $v rc=$? $(python -c "
import json; d=json.loads(open('gpurun_out/c2u_$v.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['roofline']['frac'],4))" 2>&1 | tail -1) t=$SECONDS"
  done
fi
