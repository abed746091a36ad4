#!/bin/bash
# Runs every BASELINE workload once (short) and collects the JSON lines.  Usage: bash profiles/run_all_workloads.sh OUT.jsonl
OUT=${1:-gpurun_out/all_workloads.jsonl}
: > $OUT
for spec in "cfg1_shortest_path 2000" "cfg2_longest_path 2000" "cfg3_mst 500" "cfg4_tsp_p1 500" "cfg4_mis 1000" "cfg4_tsp_p2 20" "cfg5_multicast 100" "cfg5_distcenter 100" "densest 500"; do
  set -- $spec
  echo "== $1" >&2
  timeout 900 python bench.py --workload $1 --steps $2 --warmup 3 --cpu-seconds 4 --e2e-steps 20 2>gpurun_out/err_$1.log | tail -1 >> $OUT || echo "{\"workload\": \"$1\", \"failed\": true}" >> $OUT
  tail -3 gpurun_out/err_$1.log >&2
done
