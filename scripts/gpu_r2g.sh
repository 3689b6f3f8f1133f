#!/bin/bash
# Round 2, call G: fresh-cache effect narrowed down (cache-create race fixed, graph-reset knob), attention exp2 A/B, megakernel
# gamma-in-smem knob, ncu --set full captures summarised to TEXT on the box (the reports themselves exceed gpurun's 64 MiB cap).
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
echo "=== decode A/B (B=32)"
timeout 900 python scripts/decode_ab.py --batches 32 --variants "REPS=4;FRESHKV=1,REPS=4;B2_KV_RESET_GRAPH=1,REPS=4;SLEEP=3,REPS=3" --out gpurun_out/r2g_decode_ab.jsonl > gpurun_out/r2g_decode_ab.log 2>&1; grep '^{' gpurun_out/r2g_decode_ab.log | cut -c1-420; tail -2 gpurun_out/r2g_decode_ab.log | cut -c1-200
echo "=== attention variants"
timeout 600 python scripts/attn_bench.py > gpurun_out/r2g_attn_bench.txt 2>&1; cat gpurun_out/r2g_attn_bench.txt | cut -c1-200
echo "=== megakernel knobs"
timeout 900 python scripts/mega_sweep.py --new 128 --reps 3 --configs "0,1,0,0;0,1,0,1;0,1,0,0;0,1,0,1" --out gpurun_out/r2g_mega_sweep.jsonl > gpurun_out/r2g_mega_sweep.log 2>&1; grep '^{' gpurun_out/r2g_mega_sweep.log | cut -c1-330
echo "=== tests touched by the attention / megakernel changes"
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "attn or attention or golden or small or 7b or incremental or mega or flash" > gpurun_out/r2g_tests.log 2>&1; tail -n 2 gpurun_out/r2g_tests.log; grep -E "^(FAILED|ERROR|E  )" gpurun_out/r2g_tests.log | head -10 | cut -c1-300
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_mega|gemm_bf16|gemm_skinny|projector_fused|flash_tc|decode_attn|rmsnorm|layernorm|rope_kv|gemv_kernel|splice|embed|sample_publish|vit_|argmax" --csv --log-file gpurun_out/r2g_launches.csv \
    python bench.py --steps 1 --warmup 1 --new 32 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2g_ncu_list.log 2>&1; echo "rc=$?"
python scripts/launch_shares.py gpurun_out/r2g_launches.csv | head -n 30 > gpurun_out/r2g_launch_shares.txt; gzip -f gpurun_out/r2g_launches.csv
echo "=== ncu --set full: ViT tail + projector + first prefill layers"
B2_ENCODE_GRAPH=0 timeout 1200 ncu --set full --clock-control none -k regex:"gemm_bf16|projector_fused|flash_tc|rmsnorm_row|layernorm|rope_kv" -s 150 -c 30 -o gpurun_out/prof_vit_prefill \
    python bench.py --steps 1 --warmup 0 --new 2 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2g_ncu_a.log 2>&1; echo "rc=$?"
echo "=== ncu --set full: decode megakernel"
timeout 900 ncu --set full --clock-control none -k regex:"decode_mega" -s 2 -c 2 -o gpurun_out/prof_mega \
    python bench.py --steps 1 --warmup 0 --new 6 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2g_ncu_b.log 2>&1; echo "rc=$?"
echo "=== ncu --set full: batched decode (B=32)"
timeout 1200 ncu --set full --clock-control none -k regex:"gemm_skinny|decode_attn|sample_publish|embed_tokens" -s 30 -c 12 -o gpurun_out/prof_decode_b32 \
    python bench.py --batch 32 --prompt 16 --steps 1 --warmup 0 --new 4 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2g_ncu_c.log 2>&1; echo "rc=$?"
python scripts/summarize_profiles.py r2g gpurun_out > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | head -40; du -sh gpurun_out
