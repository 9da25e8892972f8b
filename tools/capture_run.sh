#!/bin/bash
# Evidence capture on a gpurun box: the driver's bench command, the bench lines of every BASELINE workload, a K = 1000
# line (a whole TimeLimit episode: the auto-reset fires inside the timed loop), the reference arm, the ncu launch list
# and one `ncu --set full` capture of a step. Output: gpurun_out/<run>_*; then, here,  python tools/refresh_profiles.py <run>
run=${1:-r2x}
out=gpurun_out
mkdir -p $out
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${run}_bench.log 2>&1
for w in Ant4Rooms-v0:65536 AntPush-v0:32768 PointUMaze-v0:4096 SwimmerUMaze-v0:65536 PointPush-v0:65536; do
  timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > $out/${run}_bench_$(echo $w | tr ':' '_').log 2>&1
done
timeout 400 python bench.py --steps 1000 --warmup 5 --no-cpu-baseline > $out/${run}_bench_k1000.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${run}_bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${run}_launches.csv \
  python bench.py --steps 6 --warmup 6 --no-cpu-baseline > $out/${run}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maze_hkernel -s 10 -c 1 -f -o $out/${run}_prof \
  python bench.py --steps 6 --warmup 6 --no-cpu-baseline > $out/${run}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maze_hkernel -s 10 -c 1 -f -o $out/${run}_prof_point \
  python bench.py --steps 6 --warmup 6 --no-cpu-baseline --workload PointUMaze-v0:4096 > $out/${run}_ncu_full_point.log 2>&1
timeout 600 python -m pytest tests -m gpu -q > $out/${run}_pytest.log 2>&1
tail -n 1 $out/${run}_bench*.log $out/${run}_pytest.log | cut -c1-400
ls -la $out/${run}_*
