#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 1 2 4 8; do
  echo "=== bench N=$n ==="
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 1000 --warmup 100 --no-cpu 2>&1 | grep -v "Warning\|warn" | tail -1 | cut -c1-330 | tee gpurun_out/s_bench_$n.log
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 1000 --warmup 100 2>&1 | grep -v "Warning\|warn\|\*\*\*" | tail -2 | cut -c1-330 | tee gpurun_out/s_bench_$n.log
  fi
done
echo "=== eager allreduce N=8 ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 1000 --warmup 100 --eager-allreduce 2>&1 | grep -v "Warning\|warn\|\*\*\*" | tail -1 | cut -c1-330
