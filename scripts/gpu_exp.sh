#!/bin/bash
# experiment call: skinny GEMM parity + per-op timings, megakernel knob sweep, batch ladder with skinny on/off
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
rm -f gpurun_out/op_bench.jsonl gpurun_out/mega_sweep.jsonl
timeout 400 python -m pytest tests/test_ops_gpu.py -q --tb=short -x -p no:cacheprovider -k "skinny" > gpurun_out/skinny.log 2>&1; echo "skinny tests rc=$? $(tail -n 1 gpurun_out/skinny.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/skinny.log | head -20
timeout 300 python scripts/op_bench.py --batch 32,16 > gpurun_out/op_bench.log 2>&1; echo "op_bench rc=$?"; tail -n 30 gpurun_out/op_bench.log
timeout 700 python scripts/mega_sweep.py --batches 16,32 > gpurun_out/mega_sweep.log 2>&1; echo "mega_sweep rc=$?"; tail -n 16 gpurun_out/mega_sweep.log
