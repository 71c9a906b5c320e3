#!/bin/bash
# one full ncu capture of each kernel of the split step (bench workload, after settling)
mkdir -p gpurun_out
SKIP=${SKIP:-1510}
ncu --set full --import-source on --clock-control none --launch-skip $SKIP -c 5 -f -o gpurun_out/r02_split_full \
   python tools/bench_config.py models/humanoid.mjb 0 4096 20 300 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/r02_split_full.ncu-rep --page raw --csv > gpurun_out/r02_split_full_raw.csv 2>/dev/null
ls -la gpurun_out/
