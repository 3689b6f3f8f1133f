#!/bin/bash
# Round 2, call E: programmatic dependent launch on the batched decode step (A/B), time evolution of a decode run, ops tests.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
: > gpurun_out/r2e_summary.txt
for f in test_generate_gpu test_model_gpu test_ops_gpu test_fp8_gpu; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2e_$f.log 2>&1
  echo "rc=$? $f: $(tail -n 1 gpurun_out/r2e_$f.log)" | tee -a gpurun_out/r2e_summary.txt
  grep -E "^(FAILED|ERROR|E  )" gpurun_out/r2e_$f.log | head -30 | cut -c1-300 | tee -a gpurun_out/r2e_summary.txt
done
echo "=== decode A/B"
timeout 1200 python scripts/decode_ab.py --batches 8,32 --variants "default;B2_PDL=0;SLEEP=0;SLEEP=0,B2_PDL=0" --out gpurun_out/r2e_decode_ab.jsonl > gpurun_out/r2e_decode_ab.log 2>&1; echo "ab rc=$?"; grep '^{' gpurun_out/r2e_decode_ab.log | cut -c1-700
tail -3 gpurun_out/r2e_decode_ab.log | cut -c1-300
