#!/bin/bash
# experiment call 3: attention v3 (single pass, lazy rescale, 1 vs 2 CTAs/SM at d=64), vectorised RoPE, row-per-CTA RMSNorm
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 400 python -m pytest tests/test_ops_gpu.py -q --tb=short -p no:cacheprovider -k "flash or attention or rope or rmsnorm or layernorm" > gpurun_out/ops3.log 2>&1; echo "ops rc=$? $(tail -n 1 gpurun_out/ops3.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/ops3.log | head -20
timeout 300 python scripts/attn_bench.py > gpurun_out/attn_v3.log 2>&1; echo "attn_bench rc=$?"; cat gpurun_out/attn_v3.log | tail -n 24
timeout 600 python -m pytest tests/test_model_gpu.py -q --tb=short -x -p no:cacheprovider > gpurun_out/model3.log 2>&1; echo "model rc=$? $(tail -n 1 gpurun_out/model3.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/model3.log | head -20
rm -f gpurun_out/config_sweep.jsonl
timeout 300 python scripts/config_sweep.py --model 7b --vit 1,16 --prefill 1,8 --decode "" > gpurun_out/sweep3.log 2> gpurun_out/sweep3.err; echo "sweep rc=$?"; tail -n 3 gpurun_out/sweep3.err; cat gpurun_out/sweep3.log
