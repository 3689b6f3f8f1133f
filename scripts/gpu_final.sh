#!/bin/bash
# What the driver runs at round end (single-process GPU suite, smoke, N=1 bench) + the profiles committed under profiles/:
# ncu launch list of the bench command, ncu --set full of the dominant kernels, generate() stage timing.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
T0=$(date +%s)
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$? $(grep -E 'passed|failed' gpurun_out/pytest_gpu_all.log | tail -n 1) wall=$(( $(date +%s) - T0 ))s"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu_all.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -n 1 gpurun_out/smoke.log)"
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
echo "=== generate() stages"; B2_PROFILE_GENERATE=1 timeout 600 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/bench_stages.json 2> gpurun_out/bench_stages.err; echo "rc=$?"; grep "generate stages" gpurun_out/bench_stages.err | tail -n 2
echo "=== ncu launch list (same command; 1 warm-up + 1 timed step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "=== ncu full: decode megakernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 2 -c 2 -o gpurun_out/prof_mega \
    python bench.py --steps 1 --warmup 0 --new 6 --no-e2e --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/ncu_mega.log
echo "=== ncu full: batched decode (B=32): swap-AB stream-K GEMM + split-KV attention"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny|decode_attn" -s 40 -c 6 -o gpurun_out/prof_skinny \
    python bench.py --batch 32 --steps 1 --warmup 0 --new 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_skinny.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/ncu_skinny.log
echo "=== ncu full: prefill GEMM + tcgen05 attention"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05|flash_tc" -s 150 -c 6 -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 0 --new 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/ncu_gemm.log
ls -la gpurun_out/ | head -40
