#!/bin/bash
# Profiles committed under profiles/: ncu launch list of the bench command (our kernels + torch elementwise), ncu --set full
# of the dominant kernels. Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
echo "=== ncu launch list (bench command, 1 timed step; -k takes BASE kernel names and keeps the weight-init kernels of the synthetic model out)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_mega|gemm_bf16_tcgen05|gemm_skinny|flash_tc|flash_fwd|decode_attn|rmsnorm|layernorm|rope_kv|gemv_kernel|splice|embed|argmax|vit_|im2col|elementwise" --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
python scripts/launch_shares.py gpurun_out/launches.csv | head -n 24
echo "=== ncu full: decode megakernel + prefill GEMM + tcgen05 attention (one process)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_mega|gemm_bf16_tcgen05|flash_tc" -s 150 -c 8 -o gpurun_out/prof_main \
    python bench.py --steps 1 --warmup 0 --new 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_main.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/ncu_main.log
echo "=== ncu full: batched decode (B=32): swap-AB stream-K GEMM + split-KV attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny|decode_attn" -s 40 -c 6 -o gpurun_out/prof_skinny \
    python bench.py --batch 32 --prompt 16 --steps 1 --warmup 0 --new 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_skinny.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/ncu_skinny.log
ls -la gpurun_out/ | head -40
