#!/bin/bash
# 2-GPU check of the driver's launch line (run with: gpurun --gpus 2 -- 'bash scripts/gpu_r2h_2gpu.sh')
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/llava-plus-codebase_b200:$PYTHONPATH
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "n2 rc=$?"; tail -c 1500 gpurun_out/r2h_bench_n2.json; tail -n 5 gpurun_out/r2h_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2h_ref_n2.json 2> gpurun_out/r2h_ref_n2.err; echo "ref n2 rc=$?"; tail -c 600 gpurun_out/r2h_ref_n2.json
timeout 600 python -m pytest tests/test_replicas_gloo.py -q -p no:cacheprovider 2>&1 | tail -2
