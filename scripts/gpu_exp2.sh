#!/bin/bash
# experiment call 2: tcgen05 attention v2 (2 CTAs/SM at d=64, lazy rescale) + BN=192 GEMM tile: parity, timings, ViT/prefill effect
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 400 python -m pytest tests/test_ops_gpu.py -q --tb=short -p no:cacheprovider -k "flash or attention or gemm" > gpurun_out/ops_attn_gemm.log 2>&1; echo "ops(attn,gemm) rc=$? $(tail -n 1 gpurun_out/ops_attn_gemm.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/ops_attn_gemm.log | head -20
B2_FLASH_TC=1 timeout 200 python scripts/attn_bench.py > gpurun_out/attn_tc.log 2>&1; echo "attn_bench tc rc=$?"; cat gpurun_out/attn_tc.log | tail -n 10
timeout 300 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; echo "gemm_sweep rc=$?"; tail -n 18 gpurun_out/gemm_sweep.log
rm -f gpurun_out/config_sweep.jsonl
timeout 400 python scripts/config_sweep.py --model 7b --vit 1,16,64 --prefill 1,8 --decode 16,32 > gpurun_out/sweep2.log 2> gpurun_out/sweep2.err; echo "sweep rc=$?"; tail -n 3 gpurun_out/sweep2.err; cat gpurun_out/sweep2.log
