#!/bin/bash
# experiment call 5: per-kernel times inside the batched decode step (ncu launch list), B=4 and B=32
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
for B in 4 32; do
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b$B.csv \
    python bench.py --batch $B --prompt 16 --steps 1 --warmup 0 --new 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b$B.log 2>&1; echo "ncu B=$B rc=$?"
python scripts/launch_shares.py gpurun_out/launches_b$B.csv "skinny|decode_attn|rmsnorm|argmax|embed_tokens|store_token|add_i32|gemv" | head -n 14
done
B2_DECODE_SKINNY=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b4_gemv.csv \
    python bench.py --batch 4 --prompt 16 --steps 1 --warmup 0 --new 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b4_gemv.log 2>&1; echo "ncu B=4 gemv rc=$?"
python scripts/launch_shares.py gpurun_out/launches_b4_gemv.csv "skinny|decode_attn|rmsnorm|argmax|embed_tokens|store_token|add_i32|gemv" | head -n 14
