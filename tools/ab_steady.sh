#!/bin/bash
# A/B inside ONE gpurun call: base build against the in-tree build, first steps after a reset and steady state (steps 800..1000).
#   tools/ab_steady.sh WORKLOAD:N [more workloads for the short run]
w=$1
for lib in libmmz_base.so libmmz.so libmmz_base.so libmmz.so; do
  echo "== $lib"; MMZ_LIB=$lib tools/quick_time.sh --no-tests "$@"
done
for lib in libmmz_base.so libmmz.so; do
  echo "== $lib steady state"
  MMZ_LIB=$lib timeout 200 python bench.py --no-cpu-baseline --steps 200 --warmup 800 --workload "$w" 2>&1 | tail -1 | python -c 'import sys, json; d = json.loads(sys.stdin.read()); print(d["config"]["workload"], d["ms_per_step"], d["value"])'
done
