#!/bin/bash
# Run every diagnostic stage in its own process with a timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag_gpu.txt 2>&1
for stage in "$@"; do
  echo "##### $stage"
  timeout 180 python tools/gpu_diag.py $stage > gpurun_out/diag_$stage.log 2>&1
  echo "exit=$?" >> gpurun_out/diag_$stage.log
  tail -n 60 gpurun_out/diag_$stage.log
done
