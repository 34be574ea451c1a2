#!/bin/bash
# round-1 late: rollout + AdamW rows first run, full suite, bench lines, ncu evidence (small reports)
mkdir -p gpurun_out
O=gpurun_out
echo "=== gpu suite ==="
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee $O/j_pytest.log
echo "=== bench (default) ==="
timeout 300 python bench.py 2>&1 | grep "^{" | tail -1 > $O/j_bench_fp32.json; cut -c1-330 $O/j_bench_fp32.json
echo "=== rollout bench ==="
timeout 200 python scripts/rollout_bench.py 2>&1 | grep "^{" | tee $O/j_rollout.jsonl
echo "=== sim ==="
timeout 120 python scripts/sim_only.py 2>&1 | grep "^{" | tee $O/j_sim.jsonl
echo "=== ncu sim (6 launches) ==="
timeout 150 ncu --set full --clock-control none -k regex:similarity_tc -c 6 -o $O/j_ncu_sim python scripts/sim_only.py --once > $O/j_ncu_sim.log 2>&1
echo "=== ncu launch list (warm) ==="
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 70 -c 80 --csv --log-file $O/j_launches_warm.csv python bench.py --steps 6 --warmup 3 --no-cpu --nbuf 2 > $O/j_ncu1.log 2>&1
echo "=== ncu rollout ==="
timeout 150 ncu --set full --clock-control none -k regex:rollout -c 2 -o $O/j_ncu_rollout python scripts/rollout_bench.py "11,64,3,197" --no-cpu > $O/j_ncu_rollout.log 2>&1
echo "=== bench bf16 ==="
timeout 200 python bench.py --mode bf16 --no-cpu 2>&1 | grep "^{" | tail -1 > $O/j_bench_bf16.json; cut -c1-200 $O/j_bench_bf16.json
ls -la $O | grep " j_"
