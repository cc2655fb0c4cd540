#!/bin/bash
# Round-2 ncu pass (one gpurun call): launch lists of the driver-facing bench (config 4, one GPU: batch 256) and of the
# batch-32 AR decode step (BASELINE configs[1]), then --set full captures of the two dominant kernels of that step.
# Every ncu run follows a plain run of the same command that exited 0.
set -x
B32="python tools/bench_r01_decode.py --steps 20 --warmup 3 --no-extras"
BENCH="python bench.py --steps 1 --warmup 3 --no-extras"
$B32 > gpurun_out/r02p_plain_b32.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/r02_launches_ar_decode_b32.csv $B32 > gpurun_out/r02p_ncu_b32.log 2>&1
$B32 > gpurun_out/r02p_plain_b32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_gemm_kernel -s 60 -c 4 -o gpurun_out/r02_prof_decode_gemm $B32 > gpurun_out/r02p_ncu_dg.log 2>&1
$B32 > gpurun_out/r02p_plain_b32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_decode_mma -s 40 -c 2 -o gpurun_out/r02_prof_attn_decode $B32 > gpurun_out/r02p_ncu_attn.log 2>&1
$BENCH > gpurun_out/r02p_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 320 --csv --log-file gpurun_out/r02_launches_bench_1gpu.csv $BENCH > gpurun_out/r02p_ncu_bench.log 2>&1
ls -la gpurun_out/ | grep r02_
tail -3 gpurun_out/r02p_ncu_b32.log gpurun_out/r02p_ncu_dg.log gpurun_out/r02p_ncu_bench.log
