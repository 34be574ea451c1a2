#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite ==="
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 | tee gpurun_out/b_pytest.log
echo "=== tc debug ==="
timeout 300 python scripts/tc_debug.py 2>&1 | tail -14 | tee gpurun_out/b_tc_debug.log
echo "=== op times ==="
timeout 300 python scripts/op_times.py cub_b64 fp32 2>&1 | tail -22 | tee gpurun_out/b_op_times.log
echo "=== ncu full: similarity_tc (fp32 then bf16) ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:similarity_tc -s 2 -c 2 -o gpurun_out/prof_sim_fp32 python scripts/op_times.py cub_b64 fp32 > gpurun_out/b_ncu1.log 2>&1
tail -2 gpurun_out/b_ncu1.log
