#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
B2_SKINNY_TRACE=gpurun_out/r2n_sk_trace_b32.txt timeout 600 python scripts/decode_ab.py --batches 32 --new 16 --variants "REPS=2" --out gpurun_out/r2n_trace_run.jsonl > gpurun_out/r2n_trace_run_b32.log 2>&1
python scripts/skinny_trace.py gpurun_out/r2n_sk_trace_b32.txt > gpurun_out/r2n_sk_trace_b32_summary.txt 2>&1; tail -8 gpurun_out/r2n_sk_trace_b32_summary.txt | cut -c1-300
gzip -f gpurun_out/r2n_sk_trace_b32.txt
