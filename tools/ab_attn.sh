#!/bin/bash
# A/B of two builds of the library in ONE GPU run (box-to-box variance is ~8 %): transformer_b200/libb200xf_base.so vs libb200xf.so
for i in 1 2; do
  echo "== base"; B200XF_LIB_PATH=$PWD/transformer_b200/libb200xf_base.so python tools/attn_bench.py quick
  echo "== new";  python tools/attn_bench.py quick
done
