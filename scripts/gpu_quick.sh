#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite ==="
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/q_pytest.log
echo "=== op times ==="
timeout 300 python scripts/op_times.py cub_b64 fp32 2>&1 | tail -24 | grep -v "^{" | tee gpurun_out/q_op_times.log
echo "=== bench ==="
timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu 2>&1 | grep -v Warning | tail -1 | cut -c1-400 | tee gpurun_out/q_bench.log
