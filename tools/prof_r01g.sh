#!/bin/bash
# final measurement pass of round 1: full bench, launch list of the default B=32 step, ncu --set full of the new backward kernels
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err || exit 1
python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/b32_short.json 2> gpurun_out/b32_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 420 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/ncu_r1g.log 2>&1
python tools/train_time.py 4 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_d -c 4 -o gpurun_out/prof_r1g_attn_bwd -f python tools/train_time.py 4 > gpurun_out/ncu_attn_bwd.log 2>&1
ncu -i gpurun_out/prof_r1g_attn_bwd.ncu-rep --page details --csv > gpurun_out/prof_r1g_attn_bwd_details.csv 2>/dev/null
tail -2 gpurun_out/ncu_attn_bwd.log
