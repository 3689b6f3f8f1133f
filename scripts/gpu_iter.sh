#!/bin/bash
# quick iteration: model tests, short bench, megakernel phase trace
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 900 python -m pytest tests/test_model_gpu.py -q --tb=short -x -p no:cacheprovider > gpurun_out/model.log 2>&1; echo "model rc=$? $(tail -n 1 gpurun_out/model.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/model.log | head -20
B2_DECODE_MEGA=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_iter_1.json 2> gpurun_out/bench_iter_1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_iter_1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['breakdown']); print(d['roofline']['frac'], d['e2e'])" ; tail -n 3 gpurun_out/bench_iter_1.err
B2_MEGA_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --new 64 > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; echo "trace rc=$?"; head -6 gpurun_out/mega_trace.txt
