#!/bin/bash
# Round 2, call D: new paths (device splice, continuous batching, fused projector, GPU preprocessing, encode graph), decode A/B.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
: > gpurun_out/r2d_summary.txt
for f in test_generate_gpu test_preprocess test_model_gpu test_checkpoint_dir test_full_depth_gpu test_ops_gpu test_fp8_gpu; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2d_$f.log 2>&1
  echo "rc=$? $f: $(tail -n 1 gpurun_out/r2d_$f.log)" | tee -a gpurun_out/r2d_summary.txt
  grep -E "^(FAILED|ERROR|E  )" gpurun_out/r2d_$f.log | head -30 | cut -c1-300 | tee -a gpurun_out/r2d_summary.txt
done
echo "=== decode A/B"
timeout 900 python scripts/decode_ab.py --batches 8,32 --variants "default;B2_SAMPLE_LEGACY=1" > gpurun_out/r2d_decode_ab.log 2>&1; echo "ab rc=$?"; grep '^{' gpurun_out/r2d_decode_ab.log | cut -c1-400
echo "=== bench (no configs)"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?" | tee -a gpurun_out/r2d_summary.txt
tail -n 3 gpurun_out/r2d_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['breakdown'],d['e2e'],d['e2e_stream'])"
B2_ENCODE_GRAPH=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2d_bench_nograph.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2d_bench_nograph.json').read().strip().splitlines()[-1]);print('no encode graph:',d['value'],d['breakdown'])"
