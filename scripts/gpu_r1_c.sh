#!/bin/bash
# rollout row first run + full suite + ncu of six similarity launches (small reports: no source import)
mkdir -p gpurun_out
O=gpurun_out
echo "=== rollout tests ==="
timeout 300 python -m pytest tests/test_rollout_gpu.py -q -m gpu -x 2>&1 | tail -15 | tee $O/i_pytest_rollout.log
echo "=== gpu suite ==="
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $O/i_pytest.log
echo "=== rollout bench ==="
timeout 200 python scripts/rollout_bench.py 2>&1 | grep "^{" | tee $O/i_rollout.jsonl
echo "=== sim plan check ==="
timeout 120 python scripts/sim_only.py 2>&1 | grep "^{" | tee $O/i_sim.jsonl
echo "=== bench ==="
timeout 300 python bench.py --no-cpu 2>&1 | grep "^{" | tail -1 > $O/i_bench.json; cut -c1-330 $O/i_bench.json
echo "=== ncu sim (6 launches) ==="
timeout 200 ncu --set full --clock-control none -k regex:similarity_tc -c 6 -o $O/i_ncu_sim python scripts/sim_only.py --once > $O/i_ncu_sim.log 2>&1
ls -la $O | grep " i_"
