#!/bin/bash
# tuning aid: C2 device-resident throughput for several small-p tier tables (cols:warps)
run() { echo "== $1"; timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --tiers "$1" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"; }
run "36:1,64:1,96:1,154:2,204:2,284:2,420:4,856:8"
run "36:1,64:1,96:1,138:1,204:2,284:2,420:4,856:8"
run "36:1,64:1,96:1,154:2,204:2,284:4,420:4,856:8"
run "36:1,64:1,96:1,154:2,204:2,284:2,420:4,856:16"
run "64:1,96:1,154:2,204:2,284:2,420:4,856:8"
run "36:1,64:1,96:1,154:2,204:2,284:2,420:8,856:8"
