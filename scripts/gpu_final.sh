#!/bin/bash
# What the driver runs at round end: single-process GPU suite, smoke, N=1 bench (both arms, the driver's step counts).
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
T0=$(date +%s)
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$? $(grep -E 'passed|failed' gpurun_out/pytest_gpu_all.log | tail -n 1) wall=$(( $(date +%s) - T0 ))s"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu_all.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -n 1 gpurun_out/smoke.log)"
echo "=== bench"; T1=$(date +%s); timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$? wall=$(( $(date +%s) - T1 ))s"; tail -c 2500 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
echo "=== reference arm"; T2=$(date +%s); timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "rc=$? wall=$(( $(date +%s) - T2 ))s"; tail -c 1500 gpurun_out/bench_ref_n1.json
