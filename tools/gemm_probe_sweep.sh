#!/bin/bash
# knock-out sweep of K1 on the four layer-GEMM shapes of ViT-L/14 (82-image micro-batch = 21074 token rows), q4_0 / bf16.
# usage: tools/gemm_probe_sweep.sh "0 1 8 16 32 64 96" > gpurun_out/probe.txt
M=${M:-21074}
for shape in "3072 1024 0" "1024 1024 0" "4096 1024 2" "1024 4096 0"; do
  for d in $1; do
    CLIP_B200_GEMM_DBG=$d python tools/gemm_probe.py $M $shape ${QT:-2} 2>&1 | tail -1
  done
done
