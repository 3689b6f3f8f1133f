#!/bin/bash
# Round 2, call R: ncu refresh for the kernels that changed after call G (stream-K GEMM: ring depth, shared-memory fix-up;
# decode attention: split factor), summarised to text on the box; launch list of a bs=32 step.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 1200 ncu --set full --clock-control none -k regex:"gemm_skinny|decode_attn|rmsnorm_row|sample_publish" -s 30 -c 14 -o gpurun_out/prof_decode_b32 \
    python bench.py --batch 32 --prompt 16 --steps 1 --warmup 0 --new 4 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2r_ncu_c.log 2>&1; echo "rc=$?"
python scripts/summarize_profiles.py r2r gpurun_out > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_skinny|decode_attn|rmsnorm|sample_publish|embed|gemm_bf16|flash_tc|rope_kv|projector_fused|splice|vit_|layernorm" --csv --log-file gpurun_out/r2r_launches_b32.csv \
    python bench.py --batch 32 --prompt 16 --steps 1 --warmup 1 --new 8 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2r_ncu_list.log 2>&1; echo "rc=$?"
python scripts/launch_shares.py gpurun_out/r2r_launches_b32.csv | head -n 30 > gpurun_out/r2r_launch_shares_b32.txt; cat gpurun_out/r2r_launch_shares_b32.txt | head -20; gzip -f gpurun_out/r2r_launches_b32.csv
ls -la gpurun_out | head; du -sh gpurun_out
