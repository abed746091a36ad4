#!/bin/bash
# round-2 GPU call 10: register-resident sampler in the incremental tree kernels + two-deep pipelined DistributionCenter search
cd $GRAFT_REPO_ROOT
S=gpurun_out/r10_status.txt; : > $S
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -k "Steiner or Multicast or Distribution or cfg3 or cfg5 or fused or golden" > gpurun_out/r10_tests.log 2>&1; echo "tests rc=$?" >> $S
for wl in cfg5_multicast cfg3_mst cfg5_distcenter; do
  timeout 600 python bench.py --workload $wl --only-headline --steps 200 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r10_bench_$wl.json 2> gpurun_out/r10_bench_$wl.err; echo "bench $wl rc=$?" >> $S
done
