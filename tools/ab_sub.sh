#!/bin/bash
# sub-batch overlap sweep -> gpurun_out/ab_sub.jsonl.  usage: ab_sub.sh "gemm sub attn_ctas" ...
out=gpurun_out/ab_sub.jsonl
: > $out
for cfg in "$@"; do
  set -- $cfg
  VALLE_B200_DECODE_GEMM=$1 VALLE_B200_SUBBATCH=$2 VALLE_B200_ATTN_CTAS=$3 timeout 300 python bench.py --steps 300 --warmup 8 --no-extras >> $out 2>> gpurun_out/ab_sub.err
  echo "$cfg" >> gpurun_out/ab_sub.cfg
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_sub.jsonl'):
    r = json.loads(l)
    print('sub %d %-28s ms/step %.4f tok/s %.0f stepfrac %.3f attn us %.2f (rows %d) e2e %.0f' % (r['config']['sub_batches'], r['config']['decode_gemm'][:28], r['ms_per_step'], r['value'], r['config']['step_hbm_frac_of_measured_peak'], r['roofline']['us_per_launch'], r['roofline']['rows_per_launch'], r['e2e']['value']))
PY
