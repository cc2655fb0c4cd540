for b in 16 64 128 256; do
  AB_BATCH=$b AB_STEPS=150 tools/ab_r02.sh gpurun_out/r02_ab_b$b.jsonl "X=0" "VALLE_B200_DECODE_GEMM=splitk"
done
AB_BATCH=1 AB_STEPS=300 tools/ab_r02.sh gpurun_out/r02_ab_b1.jsonl "X=0"
AB_BATCH=8 AB_STEPS=300 tools/ab_r02.sh gpurun_out/r02_ab_b8.jsonl "X=0" "VALLE_B200_DECODE_GEMM=tc"
AB_BATCH=12 AB_STEPS=300 tools/ab_r02.sh gpurun_out/r02_ab_b12.jsonl "X=0" "VALLE_B200_DECODE_GEMM=splitk"
