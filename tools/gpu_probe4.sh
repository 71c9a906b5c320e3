#!/bin/bash
for g in 1 4 8 16 32; do
  MJB_SPLIT=1 MJB_GROUPS=$g timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1 | sed "s/^/G=$g /"
done
