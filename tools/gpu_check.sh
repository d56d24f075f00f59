#!/bin/bash
# run each kernel bring-up group in its own process under a timeout; logs go to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for g in "$@"; do
  echo "##### $g"
  timeout 300 python tools/gpu_check.py $g 2>&1 | tee gpurun_out/check_$g.log | tail -n 60
  echo "exit: ${PIPESTATUS[0]}"
done
