#!/bin/bash
# last call of the round: N=1 bench line first, then as much of the GPU suite as the remaining budget allows
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 420 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 3300 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
T0=$(date +%s)
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$? $(grep -E 'passed|failed' gpurun_out/pytest_gpu_all.log | tail -n 1) wall=$(( $(date +%s) - T0 ))s"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu_all.log | head -20
