#!/bin/bash
# Runs every diagnostic group in its own process (a faulting kernel must not poison the others).
# Usage (on the GPU box, from the repo root): bash tools/gpu_run_all.sh [group ...]
mkdir -p gpurun_out
groups=${@:-"simt elementwise gemm_tn gemm_mn gemm_dw attention gcn0 modules model trainer"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag_gpu.txt 2>&1
for g in $groups; do
  echo "##### $g" | tee -a gpurun_out/diag_summary.txt
  timeout 600 python tools/gpu_diag.py $g > gpurun_out/diag_$g.log 2>&1
  rc=$?
  echo "exit=$rc" | tee -a gpurun_out/diag_summary.txt
  grep -E "^(FAIL|SUMMARY)|failed:" gpurun_out/diag_$g.log | head -60 | tee -a gpurun_out/diag_summary.txt
done
