#!/bin/bash
# round-2 GPU call 17: whole GPU suite at HEAD (PDL hooks, one-tile lane blocks, warp-per-env obs, obs_x), default bench line
cd $GRAFT_REPO_ROOT
S=gpurun_out/r17_status.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r17_tests.log 2>&1; echo "tests rc=$?" >> $S
python bench.py > gpurun_out/r17_bench_default.json 2> gpurun_out/r17_bench_default.err; echo "bench rc=$?" >> $S
