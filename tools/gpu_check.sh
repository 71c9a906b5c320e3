#!/bin/bash
# quick device check used during development: GPU test suite, then the bench workload through tools/bench_config.py
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1; done
