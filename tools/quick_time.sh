#!/bin/bash
# Quick timing of a few workloads (ms per step) followed by the GPU test suite; used for A/B runs on a gpurun box.
#   tools/quick_time.sh [--no-tests] WORKLOAD:N ...
# Every command runs under its own short timeout: a hung kernel must not eat the gpurun limit.
tests=1
if [ "$1" == "--no-tests" ]; then tests=0; shift; fi
for w in "$@"; do
  timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --workload "$w" 2>&1 | tail -1 |
    python -c 'import sys, json
s = sys.stdin.read()
try:
    d = json.loads(s); print(d["config"]["workload"], d["ms_per_step"], d["value"])
except Exception:
    print("bench failed:", s[-300:]); sys.exit(3)' || exit 3
done
if [ $tests == 1 ]; then timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
