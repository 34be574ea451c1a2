#!/bin/bash
# A/B of the round-1 late variants on one GPU: similarity epilogue (PPH_SIM_EPI), programmatic dependent launch
# (PPH_PDL), stream schedule (PPH_SCHEDULE).  Every variant runs in its own process under its own timeout.
mkdir -p gpurun_out
O=gpurun_out
echo "=== gpu suite (defaults) ==="
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee $O/g_pytest.log
echo "=== gpu suite subset, PDL on ==="
PPH_PDL=1 timeout 300 python -m pytest tests/test_fused_step_gpu.py tests/test_module_gpu.py -q -m gpu -x 2>&1 | tail -8 | tee $O/g_pytest_pdl.log
CF="64,81,2000,192;256,81,2000,192;1024,81,2000,192"
echo "=== sweep: old epilogue ==="
PPH_SIM_EPI=0 timeout 200 python scripts/sweep.py "$CF" 2>&1 | grep "^{" | tee $O/g_sweep_epi0.jsonl
echo "=== sweep: new epilogue ==="
timeout 200 python scripts/sweep.py "$CF" 2>&1 | grep "^{" | tee $O/g_sweep_epi1.jsonl
echo "=== sweep: PDL ==="
PPH_PDL=1 timeout 200 python scripts/sweep.py "$CF" 2>&1 | grep "^{" | tee $O/g_sweep_pdl.jsonl
echo "=== sweep: schedule 1 ==="
PPH_SCHEDULE=1 timeout 200 python scripts/sweep.py "$CF" 2>&1 | grep "^{" | tee $O/g_sweep_sched1.jsonl
echo "=== sweep: PDL + schedule 1 ==="
PPH_PDL=1 PPH_SCHEDULE=1 timeout 200 python scripts/sweep.py "$CF" 2>&1 | grep "^{" | tee $O/g_sweep_pdl_sched1.jsonl
echo "=== bench: defaults / PDL+sched1 ==="
timeout 300 python bench.py --no-cpu 2>&1 | grep "^{" | tail -1 > $O/g_bench_default.json; cut -c1-330 $O/g_bench_default.json
PPH_PDL=1 PPH_SCHEDULE=1 timeout 300 python bench.py --no-cpu 2>&1 | grep "^{" | tail -1 > $O/g_bench_pdl_sched1.json; cut -c1-330 $O/g_bench_pdl_sched1.json
