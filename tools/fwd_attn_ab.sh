#!/bin/bash
# forward-attention A/B: parity tests, then the kernel micro-benchmark for the product build and the exp2-offload variants
python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" 2>&1 | tail -3
echo "== default (poly 2)"; python tools/kbench.py attn 2>&1 | grep '"tc": true'
for p in 0 4; do echo "== poly $p"; VALLE_B200_LIB=$PWD/valle2_b200/lib/libvalle_b200_poly$p.so python tools/kbench.py attn 2>&1 | grep '"tc": true'; done
python tools/fwd_attn_timeline.py
