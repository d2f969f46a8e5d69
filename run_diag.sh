#!/bin/bash
# Runs each diagnostic group in its own process under a timeout (a hung kernel must not hang the box).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
for g in "$@"; do
  timeout 180 python tests/diag_kernels.py $g 2>&1 | tee gpurun_out/diag_$g.log | tail -60
  echo "[exit $g: ${PIPESTATUS[0]}]"
done
