#!/bin/bash
# batch ladder / ViT throughput / 13B numbers for DESIGN.md (not the headline bench line)
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
rm -f gpurun_out/config_sweep.jsonl
timeout 700 python scripts/config_sweep.py --model 7b > gpurun_out/sweep_7b.log 2> gpurun_out/sweep_7b.err; echo "sweep 7b rc=$?"; tail -n 3 gpurun_out/sweep_7b.err
timeout 600 python scripts/config_sweep.py --model 13b --vit "" --prefill 1,4 --decode 1,4,32 > gpurun_out/sweep_13b.log 2> gpurun_out/sweep_13b.err; echo "sweep 13b rc=$?"; tail -n 3 gpurun_out/sweep_13b.err
cat gpurun_out/config_sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items()})
"
