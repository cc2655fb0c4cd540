#!/bin/bash
# A/B of the decode GEMM forms; results -> gpurun_out/ab_rows.jsonl
out=gpurun_out/ab_rows.jsonl
: > $out
: > gpurun_out/ab_rows.err
for cfg in "rows 1 1" "rows 2 1" "splitk 1 1" "rows 1 2" "$@"; do
  set -- $cfg
  echo "== gemm=$1 qkvsplit=$2 sub=$3" >> gpurun_out/ab_rows.err
  VALLE_B200_DECODE_GEMM=$1 VALLE_B200_ROWS_QKV_SPLIT=$2 VALLE_B200_SUBBATCH=$3 timeout 300 python bench.py --steps 300 --warmup 8 --no-extras >> $out 2>> gpurun_out/ab_rows.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_rows.jsonl'):
    r = json.loads(l)
    print('ms/step %.4f tok/s %.0f stepfrac %.3f sub %d gemm %s attn us %.2f frac %.3f e2e %.0f gemms ms %.3f' % (r['ms_per_step'], r['value'], r['config']['step_hbm_frac_of_measured_peak'], r['config']['sub_batches'], r['config']['decode_gemm'], r['roofline']['us_per_launch'], r['roofline']['frac'], r['e2e']['value'], r['gemm_decode']['ms_per_step']))
PY
