#!/bin/bash
# round-2 GPU call 34: ncu launch list of the headline at HEAD (streaming steps = 8 sub-batch launches each; e2e = step kernels + writeback_compact_kernel)
cd $GRAFT_REPO_ROOT
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r34_launches_default_bench.csv python bench.py --only-headline --steps 64 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/r34_launches.log 2>&1; echo "launch list rc=$?" > gpurun_out/r34_status.txt
