#!/bin/bash
for kb in 99 160; do
echo "--- prof, pool $kb KB"
MJB_ROLLOUT_POOL_KB=$kb timeout 300 python tools/prof_rollout.py 2>&1 | tail -8
done
