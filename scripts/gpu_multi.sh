#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 2 4; do
  echo "=== bench N=$n ==="
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 500 --warmup 50 2>&1 | grep -v "Warning\|warn" | tail -3 | cut -c1-700 | tee gpurun_out/m_bench_$n.log
done
echo "=== reference arm under torchrun N=2 ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 2>&1 | tail -2 | cut -c1-400
