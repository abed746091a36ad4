#!/bin/bash
# round-2 GPU call 2: full tests, e2e breakdown, A/B of the dedicated DistributionCenter kernel, ncu captures
cd $GRAFT_REPO_ROOT
S=gpurun_out/r2_status.txt; : > $S
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 300 python profiles/e2e_breakdown.py > gpurun_out/r2_e2e_breakdown.json 2> gpurun_out/r2_e2e_breakdown.err; echo "e2e rc=$?" >> $S
for wl in cfg5_distcenter cfg5_multicast cfg3_mst cfg4_mis; do
  timeout 600 python bench.py --workload $wl --only-headline --steps 100 --no-cpu --e2e-steps 10 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; echo "bench $wl rc=$?" >> $S
done
GE_NO_DC=1 timeout 600 python bench.py --workload cfg5_distcenter --only-headline --steps 100 --no-cpu --e2e-steps 10 > gpurun_out/r2_bench_cfg5_distcenter_general.json 2> gpurun_out/r2_bench_cfg5_distcenter_general.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:lane_step -c 3 -f -o gpurun_out/r2_ncu_cfg2 python bench.py --only-headline --steps 8 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r2_ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?" >> $S
timeout 600 $NCU -k regex:features_cta -c 1 -f -o gpurun_out/r2_ncu_features_n500 python profiles/feature_profile.py cfg5_multicast 2048 > gpurun_out/r2_ncu_features.log 2>&1; echo "ncu features rc=$?" >> $S
timeout 900 $NCU -k regex:dc_step -c 2 -f -o gpurun_out/r2_ncu_dc python bench.py --workload cfg5_distcenter --only-headline --steps 4 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r2_ncu_dc.log 2>&1; echo "ncu dc rc=$?" >> $S
timeout 900 $NCU -k regex:incr_tree_step -c 2 -f -o gpurun_out/r2_ncu_multicast python bench.py --workload cfg5_multicast --only-headline --steps 4 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r2_ncu_multicast.log 2>&1; echo "ncu multicast rc=$?" >> $S
ls -la gpurun_out/*.ncu-rep >> $S 2>&1
