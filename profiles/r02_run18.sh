#!/bin/bash
# round-2 GPU call 18: DistributionCenter fixed-stride rows (dc_rows): parity, A/B at config 5
cd $GRAFT_REPO_ROOT
S=gpurun_out/r18_status.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q -x -k "Distribution or distribution or dc_ or full_size or pool" > gpurun_out/r18_tests.log 2>&1; echo "tests rc=$?" >> $S
for rows in 1 0; do
  GE_DC_ROWS=$rows python bench.py --workload cfg5_distcenter --only-headline --no-cpu --no-e2e-obs --steps 200 --e2e-steps 5 > gpurun_out/r18_bench_dc_rows$rows.json 2>> gpurun_out/r18_err.log; echo "bench rows=$rows rc=$?" >> $S
done
