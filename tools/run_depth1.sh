export VALLE_B200_LIB=$PWD/valle2_b200/lib/libvalle_b200_depth1.so
tools/ab_sub.sh "rows 2 128" "rows 2 192" "rows 2 320" "srrrs 2 256" "rrrrs 2 256" "srrss 2 256" "rsrrr 2 256" "rrsrr 2 256" "rrrsr 2 256"
VALLE_B200_ROWS_QKV_SPLIT=2 tools/ab_sub.sh "rows 2 256" | tail -1
