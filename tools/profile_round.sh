#!/bin/bash
# Round profile recipe (run under gpurun on ONE B200):  bash tools/profile_round.sh r01
#   1. launch list of the bench command   -> gpurun_out/<tag>_launches.csv
#   2. one `--set full` capture of the fused step kernel inside the timed window -> gpurun_out/<tag>_step_full.ncu-rep
#      + its raw page as CSV (read here without the GUI)
#   3. the bench line itself (not under a profiler) and the reference arm
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_warp -s 315 -c 1 -f -o gpurun_out/${TAG}_step_full \
    python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_step_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_full_raw.csv 2>/dev/null
tail -1 gpurun_out/${TAG}_bench.json | cut -c1-300
