#!/bin/bash
mkdir -p gpurun_out
echo "=== bench N=2 (90 s cap) ==="
timeout -k 5 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 500 --warmup 50 2>&1 | grep -v "Warning\|warn\|\*\*\*" | tail -2 | cut -c1-400 | tee gpurun_out/s2_bench_2.log
