#!/bin/bash
# round-2 GPU call 5: DistributionCenter prefix search (tests, bench, ncu), full default bench line, features CSR profile
cd $GRAFT_REPO_ROOT
S=gpurun_out/r5_status.txt; : > $S
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -k "Distribution or cfg5 or golden or pool or properties" > gpurun_out/r5_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 600 python bench.py --workload cfg5_distcenter --only-headline --steps 100 --no-cpu --e2e-steps 10 --no-e2e-obs > gpurun_out/r5_bench_dc.json 2> gpurun_out/r5_bench_dc.err; echo "bench dc rc=$?" >> $S
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:dc_step -c 2 -f -o gpurun_out/r5_ncu_dc python bench.py --workload cfg5_distcenter --only-headline --steps 4 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r5_ncu_dc.log 2>&1; echo "ncu dc rc=$?" >> $S
timeout 600 $NCU -k regex:features_cta -c 1 -f -o gpurun_out/r5_ncu_features_csr python profiles/feature_profile.py cfg5_multicast 2048 > gpurun_out/r5_ncu_features.log 2>&1; echo "ncu features rc=$?" >> $S
start=$(date +%s)
timeout 1500 python bench.py --steps 200 > gpurun_out/r5_bench_full.json 2> gpurun_out/r5_bench_full.err; echo "bench full rc=$? wall=$(( $(date +%s) - start ))s" >> $S
