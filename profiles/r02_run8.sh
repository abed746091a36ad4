#!/bin/bash
# round-2 GPU call 8: full suite incl. PerishableProductDelivery, then the default bench line
cd $GRAFT_REPO_ROOT
S=gpurun_out/r8_status.txt; : > $S
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r8_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r8_smoke.log 2>&1; echo "smoke rc=$?" >> $S
start=$(date +%s)
timeout 1500 python bench.py --steps 200 > gpurun_out/r8_bench_full.json 2> gpurun_out/r8_bench_full.err; echo "bench full rc=$? wall=$(( $(date +%s) - start ))s" >> $S
