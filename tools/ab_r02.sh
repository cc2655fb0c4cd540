#!/bin/bash
# A/B sweep of the decode step under experiment switches: tools/ab_r02.sh OUTFILE "ENV1" "ENV2" ...   (one bench line per ENV)
out=$1; shift
: > $out
for cfg in "$@"; do
  line=$(env $cfg python tools/bench_r01_decode.py --steps ${AB_STEPS:-300} --warmup 8 --no-extras --batch ${AB_BATCH:-32} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'cfg': '$cfg', 'ms_per_step': round(d['ms_per_step'],4), 'tok_s': round(d['value']), 'hbm_frac': round(d['config']['step_hbm_frac_of_measured_peak'],3), 'e2e': round(d['e2e']['value']), 'attn_us': round(d['roofline']['us_per_launch'],2), 'gemm_ms': round(d['gemm_decode']['ms_per_step'],4), 'launches': d['config']['launches_per_step']}))")
  echo "$line" >> $out
done
cat $out
