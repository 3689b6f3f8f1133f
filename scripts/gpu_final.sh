#!/bin/bash
# what the driver runs at round end (single-process GPU suite, smoke, N=1 bench) + extra workloads for DESIGN.md
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$? $(grep -E 'passed|failed' gpurun_out/pytest_gpu_all.log | tail -n 1) $(grep -E 'Elapsed' gpurun_out/pytest_gpu_all.log)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --new 128 > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench B=32 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_b32.json')); print({k:d[k] for k in ('value','ms_per_step')}); print(d['breakdown'])"; tail -n 3 gpurun_out/bench_b32.err
timeout 900 python bench.py --model 13b --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --new 128 > gpurun_out/bench_13b.json 2> gpurun_out/bench_13b.err; echo "bench 13B rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_13b.json')); print({k:d[k] for k in ('value','ms_per_step')}); print(d['breakdown']); print(d['roofline']['frac'])"; tail -n 3 gpurun_out/bench_13b.err
