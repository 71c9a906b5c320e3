#!/bin/bash
for g in 1 2 3 4 6; do
  MJB_GROUPS=$g timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1 | sed "s/^/G=$g /"
done
MJB_GROUPS=3 timeout 600 python -m pytest tests/test_groups.py tests/test_gpu_parity.py -x -q -m gpu -k "groups or pgs or 4096" 2>&1 | tail -2
