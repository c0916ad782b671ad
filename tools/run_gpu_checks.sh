#!/bin/bash
# Runs every GPU test file in its own process (a CUDA fault in one cannot poison the others), each bounded by
# `timeout`, and writes logs + a summary under gpurun_out/checks/.  Usage: tools/run_gpu_checks.sh [files...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/checks
files=("$@")
if [ ${#files[@]} -eq 0 ]; then files=(tests/test_gpu_cc.py tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_elementwise.py tests/test_gpu_modules.py tests/test_gpu_e2e.py tests/test_gpu_etam.py tests/test_gpu_parity_r2.py); fi
: > gpurun_out/checks/summary.txt
for f in "${files[@]}"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q -s --tb=short --timeout 600 -p no:cacheprovider > "gpurun_out/checks/$name.log" 2>&1
  rc=$?
  echo "$name rc=$rc $(tail -n 1 gpurun_out/checks/$name.log)" | tee -a gpurun_out/checks/summary.txt
done
