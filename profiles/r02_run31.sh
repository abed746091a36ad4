#!/bin/bash
# round-2 GPU call 31: the default bench line and the driver-like line at HEAD (auto streamed write-back, thread-per-node obs kernel)
cd $GRAFT_REPO_ROOT
S=gpurun_out/r31_status.txt; : > $S
timeout 1500 python bench.py > gpurun_out/r31_bench_default.json 2> gpurun_out/r31_bench_default.err; echo "bench default rc=$?" >> $S
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r31_bench_driverlike.json 2> gpurun_out/r31_bench_driverlike.err; echo "bench driver-like rc=$?" >> $S
