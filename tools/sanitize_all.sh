#!/bin/bash
# compute-sanitizer logs of the step-kernel families (run on a gpurun box): gpurun_out/<run>_sanitizer_<tool>_<tag>.txt
run=${1:-r2}
out=gpurun_out
mkdir -p $out
# tag:env id:extra environment
for spec in "AntUMaze-v0:AntUMaze-v0:" "AntPush-v0:AntPush-v0:" "PointUMaze-v0:PointUMaze-v0:" \
            "PointUMaze-v0_lanes:PointUMaze-v0:MMZ_POINT_HYBRID=0" "AntMultiPush-v0:AntMultiPush-v0:" "SwimmerUMaze-v0:SwimmerUMaze-v0:"; do
  IFS=: read tag env extra <<< "$spec"
  for tool in memcheck racecheck synccheck; do
    env $extra timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $env 64 3 \
      > $out/${run}_sanitizer_${tool}_${tag}.txt 2>&1
    echo "$tag $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${run}_sanitizer_${tool}_${tag}.txt | tail -1) | $(grep -E 'x64, ' $out/${run}_sanitizer_${tool}_${tag}.txt | tail -1)"
  done
done
