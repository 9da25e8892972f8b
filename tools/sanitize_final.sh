#!/bin/bash
# compute-sanitizer on the final kernels (run on a gpurun box), all three tools on the hybrid instances, memcheck + racecheck on
# the lanes-per-environment ones: gpurun_out/<run>_sanitizer_<tool>_<tag>.txt
run=${1:-r2}
out=gpurun_out
mkdir -p $out
for spec in "AntUMaze-v0:AntUMaze-v0::memcheck racecheck synccheck" "AntPush-v0:AntPush-v0::memcheck racecheck synccheck" \
            "PointUMaze-v0:PointUMaze-v0::memcheck racecheck synccheck" "PointUMaze-v0_lanes:PointUMaze-v0:MMZ_POINT_HYBRID=0:memcheck racecheck synccheck" \
            "AntMultiPush-v0:AntMultiPush-v0::memcheck racecheck synccheck" "SwimmerUMaze-v0:SwimmerUMaze-v0::memcheck racecheck synccheck"; do
  IFS=: read tag env extra tools <<< "$spec"
  for tool in $tools; do
    t0=$(date +%s)
    env $extra timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $env 64 3 \
      > $out/${run}_sanitizer_${tool}_${tag}.txt 2>&1
    echo "$tag $tool ($(( $(date +%s) - t0 )) s): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${run}_sanitizer_${tool}_${tag}.txt | tail -1) | $(grep -E 'x64, ' $out/${run}_sanitizer_${tool}_${tag}.txt | tail -1)"
  done
done
