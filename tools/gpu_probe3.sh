#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pgs or 4096 or forward" 2>&1 | tail -5
MJB_PGS4_SLOTS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pgs or 4096" 2>&1 | tail -3
MJB_SPLIT=1 timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
MJB_PGS4_SLOTS=1 timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
bash tools/gpu_probe2.sh
