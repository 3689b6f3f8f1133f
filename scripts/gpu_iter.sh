#!/bin/bash
# quick iteration: gemv + model tests (both decode paths), short benches
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 600 python -m pytest tests/test_ops_gpu.py -q --tb=short -x -p no:cacheprovider -k "gemv" > gpurun_out/gemv.log 2>&1; echo "gemv rc=$? $(tail -n 1 gpurun_out/gemv.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/gemv.log | head -20
timeout 900 python -m pytest tests/test_model_gpu.py -q --tb=short -x -p no:cacheprovider > gpurun_out/model.log 2>&1; echo "model rc=$? $(tail -n 1 gpurun_out/model.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/model.log | head -20
B2_DECODE_MEGA=0 timeout 900 python -m pytest tests/test_model_gpu.py -q --tb=short -x -p no:cacheprovider -k "golden or 7b or small or incremental" > gpurun_out/model_mk.log 2>&1; echo "model(multi-kernel) rc=$? $(tail -n 1 gpurun_out/model_mk.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/model_mk.log | head -20
for mode in 1 0; do
B2_DECODE_MEGA=$mode timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_iter_$mode.json 2> gpurun_out/bench_iter_$mode.err; echo "bench mega=$mode rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_iter_$mode.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['breakdown']['decode_ms_per_token'], d['roofline']['frac'])" ; tail -n 3 gpurun_out/bench_iter_$mode.err
done
B2_DECODE_MEGA=0 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --batch 8 > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; echo "bench B=8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_b8.json')); print({k:d[k] for k in ('value','ms_per_step')}); print(d['breakdown'])"; tail -n 3 gpurun_out/bench_b8.err
