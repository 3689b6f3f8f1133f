#!/bin/bash
# Round 2, call C: fp8 tests after the oracle fix, full-depth parity (7B/32 layers, 13B/40 layers), checkpoint-dir ingestion,
# suites touched by the 2-CTA heuristic, bench with isolated decode timings, ncu launch list of a B=32 step.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
: > gpurun_out/r2c_summary.txt
for f in test_fp8_gpu test_checkpoint_dir test_full_depth_gpu test_model_gpu test_ops_gpu; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/r2c_$f.log 2>&1
  echo "rc=$? $f: $(tail -n 1 gpurun_out/r2c_$f.log)" | tee -a gpurun_out/r2c_summary.txt
  grep -E "^(FAILED|ERROR|E  )" gpurun_out/r2c_$f.log | head -20 | tee -a gpurun_out/r2c_summary.txt
done
cat gpurun_out/full_depth_parity.json
echo "=== bench"
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" | tee -a gpurun_out/r2c_summary.txt
tail -n 3 gpurun_out/r2c_bench.err
echo "=== ncu launch list, B=32 step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_mega|gemm_bf16|gemm_skinny|flash_tc|decode_attn|rmsnorm|layernorm|rope_kv|gemv_kernel|splice|embed|sample_publish|vit_|im2col" --csv --log-file gpurun_out/r2c_launches_b32.csv \
    python bench.py --batch 32 --steps 1 --warmup 1 --new 8 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2c_ncu_b32.log 2>&1; echo "rc=$?"
python scripts/launch_shares.py gpurun_out/r2c_launches_b32.csv | head -n 30 | tee gpurun_out/r2c_launch_shares_b32.txt
