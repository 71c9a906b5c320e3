#!/bin/bash
# development check of the persistent rollout kernel: its parity test, then the bench workload with it on
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent" 2>&1 | tail -2
for kb in 99 63 131; do
echo "--- persistent, pool $kb KB"
MJB_PERSISTENT=1 MJB_ROLLOUT_POOL_KB=$kb timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
done
