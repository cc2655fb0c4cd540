#!/bin/bash
# bench variants of the decode step; results -> gpurun_out/variants.jsonl.  usage: run_variants.sh "chain sub ctas" ...
out=gpurun_out/variants.jsonl
: > $out
for cfg in "$@"; do
  set -- $cfg
  echo "== fused=$1 sub=$2 cfg=$3" >> gpurun_out/variants.err
  VALLE_B200_FUSED=$1 VALLE_B200_SUBBATCH=$2 VALLE_B200_KV_PREFETCH="$3" timeout 300 python bench.py --steps 300 --warmup 8 --no-extras >> $out 2>> gpurun_out/variants.err
done
python - <<'PY'
import json
for l in open('gpurun_out/variants.jsonl'):
    r = json.loads(l)
    print('ms/step %.4f sub %d gemm %s attn us %.2f frac %.3f e2e %.0f gemms ms %.3f' % (r['ms_per_step'], r['config']['sub_batches'], r['config']['decode_gemm'], r['roofline']['us_per_launch'], r['roofline']['frac'], r['e2e']['value'], r['gemm_decode']['ms_per_step']))
PY
