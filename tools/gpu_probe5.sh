#!/bin/bash
for nl in 32 16 8; do
  MJB_PART_LANES=$nl timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pgs or 4096" 2>&1 | tail -1
  MJB_PART_LANES=$nl timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1 | sed "s/^/LANES=$nl /"
done
MJB_PART_LANES=8 SKIP=1510 bash tools/gpu_probe2.sh
