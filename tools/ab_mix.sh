#!/bin/bash
# decode step time for several GEMM-form mixes -> gpurun_out/ab_mix.jsonl.  usage: ab_mix.sh BATCH mix1 mix2 ...
out=gpurun_out/ab_mix.jsonl
b=$1; shift
: > $out
for m in "$@"; do
  VALLE_B200_DECODE_GEMM=$m timeout 300 python bench.py --batch $b --steps 300 --warmup 8 --no-extras >> $out 2>> gpurun_out/ab_mix.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_mix.jsonl'):
    r = json.loads(l)
    print('B %3d %-28s ms/step %.4f tok/s %.0f stepfrac %.3f attn us %.2f' % (r['config']['batch_per_gpu'], r['config']['decode_gemm'], r['ms_per_step'], r['value'], r['config']['step_hbm_frac_of_measured_peak'], r['roofline']['us_per_launch']))
PY
