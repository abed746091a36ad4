#!/bin/bash
# round-2 GPU call 14: streaming protocol sweep (rotation over R batches, C sub-batch chains, optional PDL)
cd $GRAFT_REPO_ROOT
O=gpurun_out/r14_sweep.jsonl; : > $O
S=gpurun_out/r14_status.txt; : > $S
run() { env $1 python profiles/stream_sweep.py --workload $2 --chunks $3 --steps 2000 >> $O 2>> gpurun_out/r14_err.log; echo "$1 $2 $3 rc=$?" >> $S; }
run GE_X=0 cfg2_longest_path 1,2,4,8,16
run GE_PDL=1 cfg2_longest_path 1,2,4,8
run "GE_PDL=1 GE_LANE_T=32" cfg2_longest_path 1,4
run "GE_PDL=1 GE_LANE_T=128" cfg2_longest_path 1,4
run GE_X=0 cfg1_shortest_path 1,2,4,8
run GE_PDL=1 cfg1_shortest_path 1,4
run GE_X=0 perishable 1,2,4,8
run GE_X=0 cfg3_mst 1,2,4,8
run GE_X=0 densest 1,2,4,8
GE_PDL=1 timeout 600 python -m pytest tests/test_cuda_oracle.py tests/test_cuda_golden.py -x -q -m gpu > gpurun_out/r14_tests_pdl.log 2>&1; echo "tests pdl rc=$?" >> $S
