#!/bin/bash
# tuning aid: C2 device-resident throughput for several small-p tier tables (cols:warps)
run() { echo "== $1"; timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --tiers "$1" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"; }
run "32:1,64:1,96:1,136:2,208:2,288:4,448:4,920:8"
run "32:1,64:1,96:1,136:2,208:2,288:4,448:8,920:16"
run "32:1,64:1,96:1,136:1,208:2,288:2,448:4,920:8"
run "32:1,64:1,96:2,136:2,208:4,288:4,448:8,920:16"
run "64:1,136:2,288:4,448:4,920:8"
run "48:1,96:1,160:2,288:4,448:8,920:16"
run "32:1,64:1,96:1,136:2,208:4,288:8,448:8,920:16"
