#!/bin/bash
# round-2 GPU call 12: final 1-GPU evidence -- full suite, smoke, features, Multicast capture, the default bench line and the reference arm
cd $GRAFT_REPO_ROOT
S=gpurun_out/r12_status.txt; : > $S
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r12_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r12_smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 100 python profiles/feature_profile.py cfg5_multicast 2048 > gpurun_out/r12_feat.txt 2>&1
timeout 100 python profiles/feature_profile.py cfg4_tsp_p1 2048 >> gpurun_out/r12_feat.txt 2>&1
timeout 100 python profiles/feature_profile.py cfg4_mis 2048 >> gpurun_out/r12_feat.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:incr_tree_step --launch-skip 40 -c 2 -f -o gpurun_out/r12_ncu_cfg5_multicast python bench.py --workload cfg5_multicast --only-headline --steps 64 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r12_ncu_mc.log 2>&1; echo "ncu mc rc=$?" >> $S
timeout 900 $NCU -k regex:incr_tree_step --launch-skip 40 -c 2 -f -o gpurun_out/r12_ncu_cfg3 python bench.py --workload cfg3_mst --only-headline --steps 64 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r12_ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?" >> $S
start=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r12_bench_default.json 2> gpurun_out/r12_bench_default.err; echo "bench default rc=$? wall=$(( $(date +%s) - start ))s" >> $S
start=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/r12_bench_reference.json 2> gpurun_out/r12_bench_reference.err; echo "bench reference rc=$? wall=$(( $(date +%s) - start ))s" >> $S
start=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r12_bench_driverlike.json 2> gpurun_out/r12_bench_driverlike.err; echo "bench driver-like rc=$? wall=$(( $(date +%s) - start ))s" >> $S
