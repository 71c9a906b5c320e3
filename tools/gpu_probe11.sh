#!/bin/bash
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 1500 python tools/stress_features.py 32 80 2 2>&1 | tail -8
timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
