#!/bin/bash
# Run the GPU test groups in separate processes (a CUDA fault poisons its process only), each under a timeout.
# Usage (on the GPU box, from the repo root): bash scripts/gpu_check.sh [quick]
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name timeout args...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest "$@" -q --tb=short -x -p no:cacheprovider > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "rc=$rc $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
  if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR|E  )" gpurun_out/$name.log | head -20 | tee -a gpurun_out/summary.txt; fi
}
: > gpurun_out/summary.txt
run gemm 600 tests/test_ops_gpu.py -k "gemm and not gemv"
run norms 300 tests/test_ops_gpu.py -k "layernorm or rmsnorm"
run attn 600 tests/test_ops_gpu.py -k "attention or rope"
run gemv 600 tests/test_ops_gpu.py -k "gemv"
run misc 300 tests/test_ops_gpu.py -k "argmax or im2col"
run model 1200 tests/test_model_gpu.py
B2_DECODE_MEGA=0 run model_multikernel_decode 900 tests/test_model_gpu.py -k "golden or 7b or small or incremental"
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$? $(tail -n 1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
