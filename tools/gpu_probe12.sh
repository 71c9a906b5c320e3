#!/bin/bash
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms/step %.3f e2e %.0f rollout_e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['rollout_call']['value']))"
