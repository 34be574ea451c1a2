#!/bin/bash
# rollout v2 (11/11/10-bit radix select, staged chain) against v1: parity tests for both, timing, one ncu capture
mkdir -p gpurun_out
O=gpurun_out
echo "=== rollout tests (v2 default) ==="
timeout 200 python -m pytest tests/test_rollout_gpu.py -q -m gpu 2>&1 | tail -8 | tee $O/k_pytest_v2.log
echo "=== rollout bench v2 ==="
timeout 150 python scripts/rollout_bench.py 2>&1 | grep "^{" | tee $O/k_rollout_v2.jsonl
echo "=== rollout bench v1 ==="
PPH_ROLLOUT=1 timeout 150 python scripts/rollout_bench.py "11,64,3,197" --no-cpu 2>&1 | grep "^{" | tee $O/k_rollout_v1.jsonl
echo "=== ncu rollout v2 ==="
timeout 120 ncu --set full --clock-control none -k regex:rollout -c 2 -o $O/k_ncu_rollout_v2 python scripts/rollout_bench.py "11,64,3,197" --no-cpu > $O/k_ncu_rollout.log 2>&1
ls -la $O | grep " k_"
