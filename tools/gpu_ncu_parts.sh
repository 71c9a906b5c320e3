#!/bin/bash
# development aid: full ncu capture (with source) of the two halves of one settled split step
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_step_warp --launch-skip 930 -c 3 -f \
  -o gpurun_out/parts_full python tools/bench_config.py models/humanoid.mjb 0 4096 20 300 > gpurun_out/parts_full.log 2>&1
tail -3 gpurun_out/parts_full.log
ls -la gpurun_out/parts_full.ncu-rep
