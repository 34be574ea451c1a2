#!/bin/bash
# First GPU call of round 2: the profile data the round-1 analysis is missing.
#   1. phase stamps of the four tcgemm instances (scripts/tg_debug.py)
#   2. ncu --set full --import-source of one graph-replayed step (16 kernels), small enough to pull back (< 64 MiB)
#   3. bench line + launch list for the before-state
mkdir -p gpurun_out
O=gpurun_out
#   0. the variants written after round 1's GPU budget was spent (default off): gated tests, then their timings
PPH_UNVALIDATED=1 timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $O/r2a_pytest_unvalidated.log
PPH_ROLLOUT=3 timeout 120 python scripts/rollout_bench.py "11,64,3,197;11,64,6,197" --no-cpu 2>&1 | grep "^{" | tee $O/r2a_rollout_v3.jsonl
timeout 120 python scripts/rollout_bench.py "11,64,3,197;11,64,6,197" --no-cpu 2>&1 | grep "^{" | tee $O/r2a_rollout_v2.jsonl
PPH_CLASSMAP=2 timeout 120 python scripts/next_rows_bench.py 2>&1 | grep class_maps | tee $O/r2a_classmap_v2.jsonl
timeout 120 python scripts/tg_debug.py 2>&1 | tail -12 | tee $O/r2a_tg_stamps.log
timeout 300 python bench.py --no-cpu 2>&1 | grep "^{" | tail -1 > $O/r2a_bench.json; cut -c1-300 $O/r2a_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -s 70 -c 17 -o $O/r2a_step \
    python bench.py --steps 4 --warmup 3 --no-cpu --nbuf 2 > $O/r2a_ncu.log 2>&1
ls -la $O | grep r2a_
