set -u
mkdir -p gpurun_out
B="python bench.py --config c3 --genes 592 --max-len 4000 --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 600 $B > gpurun_out/c3_592.json 2> gpurun_out/c3_592.err; echo "plain rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c3_592.json')); print(d['value'], d['roofline']['frac'], d['roofline']['buckets'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nmfoa_mid -s 6 -c 1 -o gpurun_out/prof_mid_w8 -f $B > gpurun_out/ncu_mid_w8.log 2>&1; echo "ncu rc=$?"
