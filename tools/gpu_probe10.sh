#!/bin/bash
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "islands" -s 2>&1 | grep "islands:" 
timeout 600 python bench.py --config 2 --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'value %.0f ms/step %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print(d['roofline']['kernel'], d['roofline']['frac']); print(d['cpu_baseline'])"
