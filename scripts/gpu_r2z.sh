#!/bin/bash
# Round 2, call Z: ncu --set full of the ViT tail + fused projector + first prefill layers on the final tree (the prefill QKV
# projection now carries RoPE + the cache write in its epilogue), summarised to text on the box.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
B2_ENCODE_GRAPH=0 timeout 1200 ncu --set full --clock-control none -k regex:"gemm_bf16|projector_fused|flash_tc|rmsnorm_kernel|layernorm|rope_kv" -s 150 -c 28 -o gpurun_out/prof_vit_prefill \
    python bench.py --steps 1 --warmup 0 --new 2 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2z_ncu_a.log 2>&1; echo "rc=$?"
python scripts/summarize_profiles.py r2z gpurun_out > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
grep -n "^## launch" gpurun_out/r2z_prof_vit_prefill_ncu_full.txt | cut -c1-140 | head -40
