#!/bin/bash
# compute-sanitizer logs of the four step-kernel families (run on a gpurun box): gpurun_out/<run>_sanitizer_<tool>_<env>.txt
run=${1:-r2}
out=gpurun_out
mkdir -p $out
for env in AntUMaze-v0 AntPush-v0 PointUMaze-v0 AntMultiPush-v0; do
  for tool in memcheck racecheck synccheck; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $env 64 3 \
      > $out/${run}_sanitizer_${tool}_${env}.txt 2>&1
    echo "$env $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${run}_sanitizer_${tool}_${env}.txt | tail -1) | $(grep -E 'x64, ' $out/${run}_sanitizer_${tool}_${env}.txt | tail -1)"
  done
done
