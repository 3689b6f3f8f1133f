#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
B2_MEGA_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --new 64 > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_trace.err; head -12 gpurun_out/mega_trace.txt
timeout 900 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/gemm_sweep.log | tail -20
