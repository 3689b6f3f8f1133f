#!/bin/bash
# Round 2, call L: stream-K fix-up with batched partial loads (+ residual prefetch), new ring-depth / split defaults.
mkdir -p gpurun_out
TAG=${1:-r2l}
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
show() { grep '^{' $1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['B'], d['variant'], round(d['ms_per_step'], 3), [round(x, 3) for x in d['all_ms']], round(d['frac_hbm_peak'], 3))
"; tail -1 $1 | cut -c1-150; }
echo "=== skinny / fp8 / decode tests"
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_fp8_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "skinny or fp8 or decode or batch or incremental" > gpurun_out/${TAG}_tests.log 2>&1; tail -n 2 gpurun_out/${TAG}_tests.log; grep -E "^(FAILED|ERROR|E  )" gpurun_out/${TAG}_tests.log | head -10 | cut -c1-300
echo "=== ladder"
timeout 900 python scripts/decode_ab.py --batches 8,16,32,64 --variants "REPS=3" --out gpurun_out/${TAG}_decode_ab.jsonl > gpurun_out/${TAG}_ab.log 2>&1; show gpurun_out/${TAG}_ab.log
B2_SKINNY_TRACE=gpurun_out/${TAG}_sk_trace_b32.txt timeout 600 python scripts/decode_ab.py --batches 32 --new 16 --variants "REPS=2" --out gpurun_out/${TAG}_trace_run.jsonl > gpurun_out/${TAG}_trace_run_b32.log 2>&1
python scripts/skinny_trace.py gpurun_out/${TAG}_sk_trace_b32.txt > gpurun_out/${TAG}_sk_trace_b32_summary.txt 2>&1; tail -8 gpurun_out/${TAG}_sk_trace_b32_summary.txt | cut -c1-260
rm -f gpurun_out/${TAG}_sk_trace_b32.txt
