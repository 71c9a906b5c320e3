#!/bin/bash
# round-2 evidence: GPU test suite, bench lines, ncu launch list of the bench command, one full capture of each
# kernel of the split step.  Everything lands in gpurun_out/ and is summarised into profiles/ afterwards.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_gputests.txt; cat gpurun_out/r02_gputests.txt
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_s20.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cat gpurun_out/r02_bench_reference_arm.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 1560 -c 5 -f -o gpurun_out/r02_split_full python bench.py --steps 20 --warmup 3 > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_split_full.ncu-rep --page raw --csv > gpurun_out/r02_split_full_raw.csv 2>/dev/null
cuobjdump -res-usage mujoco_b200/libmjb200.so 2>/dev/null | grep -A1 "k_pgs4\|k_step_warpILi0ELi16ELi0ELi[12]\|k_step_warpILi0ELi32ELi0ELi0" | grep -v "^--" > gpurun_out/r02_res_usage.txt
ls -la gpurun_out | tail -12
MJB_PERSISTENT=1 timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1 > gpurun_out/r02_persistent.txt; cat gpurun_out/r02_persistent.txt
timeout 300 python tools/bench_config.py models/humanoid.mjb 0 4096 200 300 2>&1 | tail -1 > gpurun_out/r02_split.txt; cat gpurun_out/r02_split.txt
