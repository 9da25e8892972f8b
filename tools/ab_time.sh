#!/bin/bash
# A/B timing inside ONE gpurun call: the in-tree libmmz.so against mujoco-maze_b200/libmmz_base.so (a copy of an earlier build).
for lib in mujoco-maze_b200/libmmz_base.so mujoco-maze_b200/libmmz.so mujoco-maze_b200/libmmz_base.so mujoco-maze_b200/libmmz.so; do
  echo "== $lib"; MMZ_LIB=$PWD/$lib tools/quick_time.sh --no-tests "$@"
done
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
