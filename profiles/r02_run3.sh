#!/bin/bash
# round-2 GPU call 3: same-box A/B of the Multicast step against the round-1 tree, e2e pipeline lanes, features CSR pull, pool tests
cd $GRAFT_REPO_ROOT
S=gpurun_out/r3_status.txt; : > $S
(cd _r01 && timeout 600 python bench.py --workload cfg5_multicast --steps 100 --no-cpu --e2e-steps 3 > ../gpurun_out/r3_r01tree_multicast.json 2> ../gpurun_out/r3_r01tree_multicast.err); echo "r01 tree multicast rc=$?" >> $S
timeout 600 python bench.py --workload cfg5_multicast --only-headline --steps 100 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r3_multicast.json 2> gpurun_out/r3_multicast.err; echo "multicast rc=$?" >> $S
timeout 300 python profiles/e2e_breakdown.py > gpurun_out/r3_e2e_breakdown.json 2> gpurun_out/r3_e2e_breakdown.err; echo "e2e rc=$?" >> $S
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -k "generate or reset_data or pipelined or heuristics or sliced" > gpurun_out/r3_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 300 python profiles/reset_costs.py 2048 > gpurun_out/r3_reset_costs.jsonl 2> gpurun_out/r3_reset_costs.err; echo "reset_costs rc=$?" >> $S
GE_FEAT_NO_CSR=1 timeout 100 python profiles/feature_profile.py cfg5_multicast 2048 > gpurun_out/r3_feat_nocsr.txt 2>&1
timeout 100 python profiles/feature_profile.py cfg5_multicast 2048 > gpurun_out/r3_feat_csr.txt 2>&1
timeout 300 python bench.py --only-headline --steps 256 --no-cpu > gpurun_out/r3_bench_cfg2.json 2> gpurun_out/r3_bench_cfg2.err; echo "bench cfg2 rc=$?" >> $S
