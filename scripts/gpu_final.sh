#!/bin/bash
# round-1 final measurements (one GPU)
mkdir -p gpurun_out
echo "=== gpu suite ==="
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/f_pytest.log
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/f_smoke.log
echo "=== bench fp32 (default) ==="
timeout 600 python bench.py 2>&1 | grep -v "Warning\|warn" | tail -1 > gpurun_out/f_bench_fp32.json; cut -c1-400 gpurun_out/f_bench_fp32.json
echo "=== bench bf16 ==="
timeout 600 python bench.py --mode bf16 --no-cpu 2>&1 | grep -v "Warning\|warn" | tail -1 > gpurun_out/f_bench_bf16.json; cut -c1-200 gpurun_out/f_bench_bf16.json
echo "=== reference arm ==="
timeout 400 python bench.py --impl reference --steps 30 --warmup 3 2>&1 | tail -1 > gpurun_out/f_bench_ref.json; cut -c1-200 gpurun_out/f_bench_ref.json
echo "=== op times ==="
timeout 300 python scripts/op_times.py cub_b64 fp32 2>&1 | grep -v "^{" | tail -18 | tee gpurun_out/f_op_times.log
echo "=== sweep ==="
timeout 500 python scripts/sweep.py 2>&1 | grep "^{" | tee gpurun_out/f_sweep.jsonl
echo "=== ncu launch list (warm caches) ==="
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 70 -c 80 --csv --log-file gpurun_out/f_launches_warm.csv python bench.py --steps 6 --warmup 3 --no-cpu --nbuf 2 > gpurun_out/f_ncu1.log 2>&1
echo "=== ncu launch list (default cache control) ==="
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 80 --csv --log-file gpurun_out/f_launches_cold.csv python bench.py --steps 6 --warmup 3 --no-cpu --nbuf 2 > gpurun_out/f_ncu2.log 2>&1
echo "=== ncu full: one launch of each kernel of the step ==="
timeout 600 ncu --set full --clock-control none --import-source on -s 70 -c 16 -o gpurun_out/f_prof_step python bench.py --steps 4 --warmup 3 --no-cpu --nbuf 2 > gpurun_out/f_ncu3.log 2>&1
ls -la gpurun_out/ | grep "f_"
