#!/bin/bash
# round-2 GPU call 9: PerishableProductDelivery lane kernel tests, final ncu captures of every workload's step kernel, launch list
cd $GRAFT_REPO_ROOT
S=gpurun_out/r9_status.txt; : > $S
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -k "Perishable or golden or properties or single" > gpurun_out/r9_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 300 python bench.py --workload perishable --only-headline --steps 200 --no-cpu --e2e-steps 10 --no-e2e-obs > gpurun_out/r9_bench_ppd.json 2> gpurun_out/r9_bench_ppd.err; echo "bench ppd rc=$?" >> $S
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name kernel-regex workload skip
  timeout 900 $NCU -k regex:$2 --launch-skip $4 -c 2 -f -o gpurun_out/r9_ncu_$1 python bench.py --workload $3 --only-headline --steps 64 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r9_ncu_$1.log 2>&1; echo "ncu $1 rc=$?" >> $S
}
cap cfg2 lane_step cfg2_longest_path 40
cap cfg1 lane_step cfg1_shortest_path 40
cap cfg3 incr_tree_step cfg3_mst 40
cap cfg4_tsp_p1 group_step cfg4_tsp_p1 40
cap cfg4_tsp_p2 group_step cfg4_tsp_p2 40
cap cfg4_mis incr_mis_step cfg4_mis 40
cap cfg5_multicast incr_tree_step cfg5_multicast 40
cap cfg5_distcenter dc_step cfg5_distcenter 40
cap densest group_step densest 40
cap perishable ppd_lane_step perishable 40
timeout 600 $NCU -k regex:features_cta -c 1 -f -o gpurun_out/r9_ncu_features_tsp200 python profiles/feature_profile.py cfg4_tsp_p1 2048 > gpurun_out/r9_ncu_features_tsp.log 2>&1; echo "ncu features tsp rc=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r9_launches_default_bench.csv python bench.py --only-headline --steps 64 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/r9_launches.log 2>&1; echo "launch list rc=$?" >> $S
ls -la gpurun_out/*.ncu-rep | wc -l >> $S
