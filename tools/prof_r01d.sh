#!/bin/bash
# round-1d measurement pass: default bench, launch list of the batch-1 lean path, ncu --set full of the rows kernel
python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err || exit 1
python bench.py --batch 1 --steps 4 --warmup 3 --no-extras > gpurun_out/b1_short.json 2> gpurun_out/b1_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 256 --csv --log-file gpurun_out/launches_r1d_b1.csv python bench.py --batch 1 --steps 4 --warmup 3 --no-extras > gpurun_out/ncu_r1d_b1.log 2>&1
python tools/prof_one.py rows_decode 1 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:linear_decode_rows -c 12 -o gpurun_out/prof_r1d_rows_b1 -f python tools/prof_one.py rows_decode 1 > gpurun_out/ncu_rows.log 2>&1
ncu -i gpurun_out/prof_r1d_rows_b1.ncu-rep --page details --csv > gpurun_out/prof_r1d_rows_b1_details.csv 2>/dev/null
tail -3 gpurun_out/ncu_rows.log
