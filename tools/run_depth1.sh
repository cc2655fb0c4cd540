# sub-batch overlap on the depth-1 PDL build vs the product build (tools/ab_sub.sh "gemm sub attn_ctas")
export VALLE_B200_LIB=$PWD/valle2_b200/lib/libvalle_b200_depth1.so
tools/ab_sub.sh "splitk 1 0" "rsrrr 2 256" "rows 2 256" "splitk 2 256" "srsss 2 256"
unset VALLE_B200_LIB
tools/ab_sub.sh "splitk 1 0" "rsrrr 2 256" "splitk 2 256"
