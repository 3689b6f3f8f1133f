#!/bin/bash
# tcgen05 flash attention bring-up: parity tests, timing vs the mma.sync kernel, then the whole GPU suite in one process
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 300 python -m pytest tests/test_ops_gpu.py -q --tb=short -p no:cacheprovider -k "flash" > gpurun_out/flash.log 2>&1; echo "flash(tc) rc=$? $(tail -n 1 gpurun_out/flash.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/flash.log | head -20
B2_FLASH_TC=1 timeout 300 python scripts/attn_bench.py > gpurun_out/attn_tc.log 2>&1; echo "attn_bench tc rc=$?"; cat gpurun_out/attn_tc.log | tail -n 12
B2_FLASH_TC=0 timeout 300 python scripts/attn_bench.py > gpurun_out/attn_mma.log 2>&1; echo "attn_bench mma rc=$?"; cat gpurun_out/attn_mma.log | tail -n 12
T0=$(date +%s)
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$? $(grep -E 'passed|failed' gpurun_out/pytest_gpu_all.log | tail -n 1) wall=$(( $(date +%s) - T0 ))s"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu_all.log | head -20
B2_MEGA_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; echo "bench(trace) rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_trace.json')); print({k:d[k] for k in ('value','ms_per_step')}); print(d['breakdown'])"; tail -n 3 gpurun_out/bench_trace.err
