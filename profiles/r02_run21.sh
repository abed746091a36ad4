#!/bin/bash
# round-2 GPU call 21: final 1-GPU evidence at HEAD -- full suite, smoke, default bench line, reference arm, driver-like line,
# ncu captures of every workload's step kernel (full-batch launches, isolated protocol), launch list of the streaming headline
cd $GRAFT_REPO_ROOT
S=gpurun_out/r21_status.txt; : > $S
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r21_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r21_smoke.log 2>&1; echo "smoke rc=$?" >> $S
start=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r21_bench_default.json 2> gpurun_out/r21_bench_default.err; echo "bench default rc=$? wall=$(( $(date +%s) - start ))s" >> $S
timeout 600 python bench.py --impl reference > gpurun_out/r21_bench_reference.json 2> gpurun_out/r21_bench_reference.err; echo "bench reference rc=$?" >> $S
start=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r21_bench_driverlike.json 2> gpurun_out/r21_bench_driverlike.err; echo "bench driver-like rc=$? wall=$(( $(date +%s) - start ))s" >> $S
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name kernel-regex workload skip
  timeout 900 $NCU -k regex:$2 --launch-skip $4 -c 2 -f -o gpurun_out/r21_ncu_$1 python bench.py --workload $3 --only-headline --no-streaming --steps 64 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r21_ncu_$1.log 2>&1; echo "ncu $1 rc=$?" >> $S
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
timeout 600 $NCU -k regex:obs_kernel -c 2 -f -o gpurun_out/r21_ncu_obs_cfg2 python bench.py --only-headline --no-streaming --steps 8 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/r21_ncu_obs.log 2>&1; echo "ncu obs rc=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r21_launches_default_bench.csv python bench.py --only-headline --steps 64 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/r21_launches.log 2>&1; echo "launch list rc=$?" >> $S
