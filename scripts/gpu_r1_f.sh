#!/bin/bash
# last validation of round 1: the whole GPU suite, smoke(), one short bench line
mkdir -p gpurun_out
O=gpurun_out
timeout 150 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $O/l_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/l_smoke.log
timeout 90 python bench.py --no-cpu --steps 1000 2>&1 | grep "^{" | tail -1 > $O/l_bench.json; cut -c1-250 $O/l_bench.json
