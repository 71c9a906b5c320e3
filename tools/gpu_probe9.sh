#!/bin/bash
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms/step %.3f e2e %.0f rollout_e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['rollout_call']['value'])); print(d['roofline']['launches_ms'])"
timeout 300 python tools/bench_config.py models/boxes.mjb 2 8192 100 100 2>&1 | tail -1
