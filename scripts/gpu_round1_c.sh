#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite ==="
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 | tee gpurun_out/c_pytest.log
echo "=== op times ==="
timeout 300 python scripts/op_times.py cub_b64 fp32 2>&1 | tail -24 | tee gpurun_out/c_op_times.log
echo "=== bench ==="
timeout 600 python bench.py --steps 1000 --warmup 100 2>&1 | grep -v Warning | tail -2 | tee gpurun_out/c_bench_fp32.log
