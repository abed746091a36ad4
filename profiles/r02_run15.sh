#!/bin/bash
# round-2 GPU call 15: parity of sub-batch streams + PDL, PDL sweeps for the other kernel families, the integrated bench line
cd $GRAFT_REPO_ROOT
O=gpurun_out/r15_sweep.jsonl; : > $O
S=gpurun_out/r15_status.txt; : > $S
timeout 900 python -m pytest tests/test_cuda_streams.py -x -q -m gpu > gpurun_out/r15_tests_streams.log 2>&1; echo "tests streams rc=$?" >> $S
run() { env $1 python profiles/stream_sweep.py --workload $2 --chunks $3 --steps $4 >> $O 2>> gpurun_out/r15_err.log; echo "$1 $2 $3 rc=$?" >> $S; }
run "GE_PDL=1 GE_LANE_T=32" cfg2_longest_path 8,16 2000
run "GE_PDL=1 GE_LANE_T=32" cfg1_shortest_path 1,4,8 2000
run GE_PDL=1 cfg1_shortest_path 8 2000
run GE_PDL=1 perishable 1,2,4 2000
run GE_PDL=1 cfg3_mst 1,2,4 2000
run GE_PDL=1 densest 1,2,4 2000
run GE_PDL=1 cfg5_multicast 1,2,4 400
run GE_X=0 cfg5_multicast 1 400
run GE_PDL=1 cfg5_distcenter 1,2,4 200
run GE_X=0 cfg5_distcenter 1 200
python bench.py > gpurun_out/r15_bench_default.json 2> gpurun_out/r15_bench_default.err; echo "bench rc=$?" >> $S
