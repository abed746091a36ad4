#!/bin/bash
# round-2 GPU call 4: e2e variants, full test suite, features (unrolled CSR gathers), the full default bench line
cd $GRAFT_REPO_ROOT
S=gpurun_out/r4_status.txt; : > $S
timeout 300 python profiles/e2e_breakdown.py > gpurun_out/r4_e2e_default.json 2> gpurun_out/r4_e2e.err
GE_PIPE_ZC=1 timeout 300 python profiles/e2e_breakdown.py > gpurun_out/r4_e2e_zc.json 2>> gpurun_out/r4_e2e.err
GE_HOST_SPIN=1 timeout 300 python profiles/e2e_breakdown.py > gpurun_out/r4_e2e_spin.json 2>> gpurun_out/r4_e2e.err
echo "e2e done" >> $S
timeout 100 python profiles/feature_profile.py cfg5_multicast 2048 > gpurun_out/r4_feat_csr.txt 2>&1
timeout 100 python profiles/feature_profile.py cfg3_mst 4096 >> gpurun_out/r4_feat_csr.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r4_tests.log 2>&1; echo "tests rc=$?" >> $S
/usr/bin/time -v timeout 1500 python bench.py --steps 200 > gpurun_out/r4_bench_full.json 2> gpurun_out/r4_bench_full.err; echo "bench full rc=$?" >> $S
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r4_bench_ref.json 2> gpurun_out/r4_bench_ref.err; echo "bench ref rc=$?" >> $S
