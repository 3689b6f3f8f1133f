#!/bin/bash
# Round 2, call A: run the two drafts that never executed (fp8 decode path, CTA-pair GEMM), one test per process so a CUDA fault
# in one does not poison the rest; then the ncu launch list of the bench command.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
: > gpurun_out/r2a_summary.txt
run_each() { # env-assignment file pattern
  local envv=$1; local file=$2; local pat=$3
  for id in $(env $envv python -m pytest $file -k "$pat" --collect-only -q -p no:cacheprovider 2>/dev/null | grep '::'); do
    env $envv timeout 300 python -m pytest "$id" -q --tb=short -p no:cacheprovider > gpurun_out/r2a_one.log 2>&1
    rc=$?
    echo "rc=$rc $id" | tee -a gpurun_out/r2a_summary.txt
    if [ $rc -ne 0 ]; then grep -E "^(E  |FAILED)" gpurun_out/r2a_one.log | head -12 | tee -a gpurun_out/r2a_summary.txt; fi
  done
}
run_each B2_TEST_FP8=1 tests/test_fp8_gpu.py ""
run_each B2_TEST_2CTA=1 tests/test_ops_gpu.py "2cta"
echo "=== ncu launch list (bench command, 1 timed step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_mega|gemm_bf16_tcgen05|gemm_skinny|flash_tc|flash_fwd|decode_attn|rmsnorm|layernorm|rope_kv|gemv_kernel|splice|embed|argmax|vit_|im2col" --csv --log-file gpurun_out/r2a_launches.csv \
    python bench.py --steps 1 --warmup 0 --new 32 --no-e2e --no-cpu-baseline > gpurun_out/r2a_ncu_list.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2a_launches.csv
python scripts/launch_shares.py gpurun_out/r2a_launches.csv | head -n 30 | tee gpurun_out/r2a_launch_shares.txt
