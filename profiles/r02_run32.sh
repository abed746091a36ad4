#!/bin/bash
# round-2 GPU call 32: whole GPU suite + smoke at HEAD; ncu capture of the thread-per-node observation kernel (DistributionCenter, config 5)
cd $GRAFT_REPO_ROOT
S=gpurun_out/r32_status.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r32_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r32_smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 400 ncu --set full --clock-control none --import-source on -k regex:obs_kernel -c 2 -f -o gpurun_out/r32_ncu_obs_distcenter python bench.py --workload cfg5_distcenter --only-headline --no-streaming --steps 8 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/r32_ncu_obs.log 2>&1; echo "ncu obs rc=$?" >> $S
