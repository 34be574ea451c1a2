#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite ==="
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/d_pytest.log
echo "=== ncu launch list, warm caches (fused step) ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 120 --csv --log-file gpurun_out/d_launches_warm.csv python bench.py --steps 6 --warmup 3 --no-cpu --nbuf 2 > gpurun_out/d_ncu.log 2>&1
tail -1 gpurun_out/d_ncu.log | cut -c1-300
