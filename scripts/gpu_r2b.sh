#!/bin/bash
# Round 2, call B: whole GPU suite after the streaming/sampling/KV-pool rework (one process per file), bench with the
# configs block + e2e_stream, GEMM tile sweep including the CTA-pair kernel.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
: > gpurun_out/r2b_summary.txt
for f in test_generate_gpu test_fp8_gpu test_model_gpu test_ops_gpu; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2b_$f.log 2>&1
  echo "rc=$? $f: $(tail -n 1 gpurun_out/r2b_$f.log)" | tee -a gpurun_out/r2b_summary.txt
  grep -E "^(FAILED|ERROR)" gpurun_out/r2b_$f.log | head -20 | tee -a gpurun_out/r2b_summary.txt
done
echo "=== bench (N=1, configs block)"
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" | tee -a gpurun_out/r2b_summary.txt
tail -c 6000 gpurun_out/r2b_bench.json; tail -n 5 gpurun_out/r2b_bench.err
echo "=== gemm sweep"
timeout 600 python scripts/gemm_sweep.py > gpurun_out/r2b_gemm_sweep.log 2>&1; echo "sweep rc=$?"; cp gpurun_out/gemm_sweep.json gpurun_out/r2b_gemm_sweep.json 2>/dev/null
