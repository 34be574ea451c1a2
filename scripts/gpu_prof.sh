#!/bin/bash
mkdir -p gpurun_out
echo "=== failing test ==="
timeout 600 python -m pytest tests/test_fused_step_gpu.py -q -m gpu -k bit_reproducible 2>&1 | grep -E "^E  |assert|Error" | cut -c1-250 | head -20 | tee gpurun_out/p_fail.log
echo "=== ncu full, one launch of each of our kernels inside the fused step ==="
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -s 70 -c 19 -o gpurun_out/prof_step python bench.py --steps 4 --warmup 3 --no-cpu --nbuf 2 > gpurun_out/p_ncu.log 2>&1
tail -2 gpurun_out/p_ncu.log | cut -c1-200
ls -la gpurun_out/prof_step.ncu-rep
