#!/bin/bash
# N=1 bench + ncu launch list (reduced decode length) + ncu --set full captures of the dominant kernels.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
echo "=== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_n1.json; tail -n 5 gpurun_out/bench_n1.err
echo "=== ncu launch list (same command; 1 warm-up + 1 timed step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "=== ncu full: decode megakernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 2 -c 2 -o gpurun_out/prof_mega \
    python bench.py --steps 1 --warmup 0 --new 6 --no-e2e --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/ncu_mega.log
echo "=== ncu full: gemm"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 100 -c 4 -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 0 --new 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; echo "rc=$?"
ls -la gpurun_out/ | head -30
