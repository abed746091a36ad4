#!/bin/bash
# round-2 GPU call 20: DistributionCenter search v5 (unrolled passes, cold overflow, ballot-built reach set): parity, A/B incl. transposed mask
cd $GRAFT_REPO_ROOT
S=gpurun_out/r20_status.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q -x -k "Distribution or distribution or dc_ or full_size or pool" > gpurun_out/r20_tests.log 2>&1; echo "tests rc=$?" >> $S
run() { env $1 python bench.py --workload cfg5_distcenter --only-headline --no-cpu --no-e2e-obs --steps 200 --e2e-steps 5 > gpurun_out/r20_bench_$2.json 2>> gpurun_out/r20_err.log; echo "bench $2 rc=$?" >> $S; }
run GE_X=0 v5
run GE_DC_TRANSPOSED=1 v5_transposed
run GE_DC_ROWS=0 csr
