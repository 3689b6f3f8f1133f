#!/bin/bash
# Round 2, call F: PDL on the prefill / ViT kernels (suite + bench), fresh-cache experiment, ncu evidence (launch list + --set full).
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
: > gpurun_out/r2f_summary.txt
for f in test_generate_gpu test_model_gpu test_ops_gpu test_fp8_gpu test_preprocess test_checkpoint_dir; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2f_$f.log 2>&1
  echo "rc=$? $f: $(tail -n 1 gpurun_out/r2f_$f.log)" | tee -a gpurun_out/r2f_summary.txt
  grep -E "^(FAILED|ERROR|E  )" gpurun_out/r2f_$f.log | head -30 | cut -c1-300 | tee -a gpurun_out/r2f_summary.txt
done
echo "=== bench PDL on / off"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2f_bench_pdl.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?" | tee -a gpurun_out/r2f_summary.txt
B2_PDL=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2f_bench_nopdl.json 2>/dev/null
python -c "
import json
for f in ('r2f_bench_pdl','r2f_bench_nopdl'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['breakdown'])"
echo "=== fresh-cache experiment"
timeout 900 python scripts/decode_ab.py --batches 32 --variants "REPS=6;FRESHKV=1,REPS=6" --out gpurun_out/r2f_decode_ab.jsonl > gpurun_out/r2f_decode_ab.log 2>&1; grep '^{' gpurun_out/r2f_decode_ab.log | cut -c1-600
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_mega|gemm_bf16|gemm_skinny|projector_fused|flash_tc|decode_attn|rmsnorm|layernorm|rope_kv|gemv_kernel|splice|embed|sample_publish|vit_|argmax" --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 1 --warmup 1 --new 32 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2f_ncu_list.log 2>&1; echo "rc=$?"
python scripts/launch_shares.py gpurun_out/r2f_launches.csv | head -n 30 | tee gpurun_out/r2f_launch_shares.txt
echo "=== ncu --set full: ViT tail + projector + first prefill layers"
B2_ENCODE_GRAPH=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|projector_fused|flash_tc|rmsnorm_row|layernorm|rope_kv" -s 150 -c 36 -o gpurun_out/prof_vit_prefill \
    python bench.py --steps 1 --warmup 0 --new 2 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2f_ncu_a.log 2>&1; echo "rc=$?"
echo "=== ncu --set full: decode megakernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_mega" -s 2 -c 2 -o gpurun_out/prof_mega \
    python bench.py --steps 1 --warmup 0 --new 6 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2f_ncu_b.log 2>&1; echo "rc=$?"
echo "=== ncu --set full: batched decode (B=32)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny|decode_attn|sample_publish|embed_tokens" -s 30 -c 12 -o gpurun_out/prof_decode_b32 \
    python bench.py --batch 32 --prompt 16 --steps 1 --warmup 0 --new 4 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2f_ncu_c.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
