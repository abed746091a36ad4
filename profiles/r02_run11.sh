#!/bin/bash
# round-2 GPU call 11: chunk-count sampler for large incremental masks
cd $GRAFT_REPO_ROOT
S=gpurun_out/r11_status.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -k "Steiner or Multicast or cfg3 or cfg5 or fused or golden or properties or pool or pipelined or sliced" > gpurun_out/r11_tests.log 2>&1; echo "tests rc=$?" >> $S
for wl in cfg5_multicast cfg3_mst; do
  timeout 600 python bench.py --workload $wl --only-headline --steps 200 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r11_bench_$wl.json 2> gpurun_out/r11_bench_$wl.err; echo "bench $wl rc=$?" >> $S
done
