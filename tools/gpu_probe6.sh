#!/bin/bash
export MJB_PART_LANES=16
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pgs or 4096" 2>&1 | tail -1
timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
python tools/prof_pgs4.py 2>&1 | head -12
