#!/bin/bash
# development probe: per-kernel durations of the split step (ncu launch list; times are serialised / cold-cache)
mkdir -p gpurun_out
SKIP=${SKIP:-1510}
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $SKIP -c 60 --csv --log-file gpurun_out/launches_split.csv \
   python tools/bench_config.py models/humanoid.mjb 0 4096 40 300 > gpurun_out/ncu_split.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_split.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(',', '')) * ({'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r[ui], 1)))
    except Exception: pass
for k, v in agg.items(): print('%-62s n=%3d mean %9.1f us  max %9.1f' % (k, len(v), sum(v) / len(v), max(v)))
PY
