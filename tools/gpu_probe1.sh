#!/bin/bash
# development probe (round 2): latency constants, parity of the split step, split vs fused timing
mkdir -p gpurun_out
./build/lat_probe > gpurun_out/lat_probe.txt 2>&1
cat gpurun_out/lat_probe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pgs or 4096 or forward" 2>&1 | tail -15
for split in 0 1; do
  MJB_SPLIT=$split timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
done
MJB_SPLIT=1 MJB_GROUPS=2 timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
MJB_SPLIT=1 MJB_GROUPS=4 timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
