#!/bin/bash
# experiment call 4: attention mask-free interior tiles + ex2.approx; swap-AB stream-K GEMM for batch 3..8
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 300 python -m pytest tests/test_ops_gpu.py -q --tb=short -p no:cacheprovider -k "flash or attention" > gpurun_out/ops4.log 2>&1; echo "ops rc=$? $(tail -n 1 gpurun_out/ops4.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/ops4.log | head -20
timeout 300 python scripts/attn_bench.py > gpurun_out/attn_v4.log 2>&1; echo "attn_bench rc=$?"; grep -v "mma.sync" gpurun_out/attn_v4.log | tail -n 16
timeout 600 python -m pytest tests/test_model_gpu.py -q --tb=short -x -p no:cacheprovider -k "13b or above_8 or batched or golden or small or 7b" > gpurun_out/model4.log 2>&1; echo "model rc=$? $(tail -n 1 gpurun_out/model4.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/model4.log | head -20
rm -f gpurun_out/config_sweep.jsonl
timeout 400 python scripts/config_sweep.py --model 7b --vit 16,64 --prefill "" --decode 3,4,8,16,32 > gpurun_out/sweep4.log 2> gpurun_out/sweep4.err; echo "sweep rc=$?"; tail -n 3 gpurun_out/sweep4.err; cat gpurun_out/sweep4.log
