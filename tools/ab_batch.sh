#!/bin/bash
# decode step time per batch size for the decode GEMM forms -> gpurun_out/ab_batch.jsonl.  usage: ab_batch.sh "1 4 8" "lean rows splitk"
out=gpurun_out/ab_batch.jsonl
: > $out
for b in $1; do
  for gm in $2; do
    VALLE_B200_DECODE_GEMM=$gm timeout 300 python bench.py --batch $b --steps 300 --warmup 8 --no-extras >> $out 2>> gpurun_out/ab_batch.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_batch.jsonl'):
    r = json.loads(l)
    print('B %3d %-40s ms/step %.4f tok/s %.0f stepfrac %.3f gemms ms %.3f launches %d' % (r['config']['batch_per_gpu'], r['config']['decode_gemm'][:40], r['ms_per_step'], r['value'], r['config']['step_hbm_frac_of_measured_peak'], r['gemm_decode']['ms_per_step'], r['config']['launches_per_step']))
PY
