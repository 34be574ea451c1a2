#!/bin/bash
# first GPU contact: CUDA-core kernels, then the tcgen05 kernel, then the full suite, a bench line and an ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi.txt
echo "=== stage 1: CUDA-core kernels ==="
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "fp32_fma or select or addon or materialised" 2>&1 | tail -25 | tee gpurun_out/stage1.log
echo "=== stage 2: tcgen05 debug ==="
timeout 300 python scripts/tc_debug.py 2>&1 | tail -40 | tee gpurun_out/stage2.log
echo "=== stage 3: full gpu suite ==="
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 | tee gpurun_out/stage3.log
echo "=== stage 4: bench ==="
timeout 600 python bench.py --steps 500 --warmup 50 2>&1 | tail -5 | tee gpurun_out/bench_fp32.log
timeout 600 python bench.py --steps 500 --warmup 50 --mode bf16 --no-cpu 2>&1 | tail -5 | tee gpurun_out/bench_bf16.log
echo "=== stage 5: ncu launch list ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --nbuf 2 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
