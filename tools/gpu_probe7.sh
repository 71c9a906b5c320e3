#!/bin/bash
export MJB_PART_LANES=16
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
