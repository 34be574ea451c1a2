#!/bin/bash
# Why did the similarity kernel get slower?  old library vs new, plan knobs, and one ncu capture at B=1024 / B=64.
mkdir -p gpurun_out
O=gpurun_out
OLD=protopformer_b200/lib/libprotohead_r1base.so
echo "=== round-1 base library ==="
timeout 120 python scripts/sim_only.py --lib $OLD 2>&1 | grep "^{" | tee $O/h_sim_old.jsonl
echo "=== current library, default plan ==="
timeout 120 python scripts/sim_only.py 2>&1 | grep "^{" | tee $O/h_sim_new.jsonl
echo "=== current library, dedicated global CTAs (round-1 plan) ==="
PPH_SIM_DEDICATED=1 timeout 120 python scripts/sim_only.py 2>&1 | grep "^{" | tee $O/h_sim_new_ded.jsonl
echo "=== current library, 4 lanes ==="
PPH_SIM_LANES=4 timeout 120 python scripts/sim_only.py 2>&1 | grep "^{" | tee $O/h_sim_new_l4.jsonl
echo "=== current library, dedicated + old epilogue ==="
PPH_SIM_DEDICATED=1 PPH_SIM_EPI=0 timeout 120 python scripts/sim_only.py 2>&1 | grep "^{" | tee $O/h_sim_new_ded_epi0.jsonl
echo "=== ncu: old library B=1024,64 ==="
timeout 300 ncu --set full --clock-control none --import-source on -k regex:similarity_tc -o $O/h_ncu_old python scripts/sim_only.py --lib $OLD --batches 1024,64 --once > $O/h_ncu_old.log 2>&1
echo "=== ncu: current library B=1024,64 ==="
timeout 300 ncu --set full --clock-control none --import-source on -k regex:similarity_tc -o $O/h_ncu_new python scripts/sim_only.py --batches 1024,64 --once > $O/h_ncu_new.log 2>&1
ls -la $O | grep " h_"
