#!/bin/bash
# round-2 GPU call 27: streamed host step (chunks = 0): parity, then e2e A/B against two slices
cd $GRAFT_REPO_ROOT
S=gpurun_out/r27_status.txt; : > $S
timeout 600 python -m pytest tests/test_cuda_streams.py tests/test_cuda_oracle.py -m gpu -q -x -k "streamed or pipelined" > gpurun_out/r27_tests.log 2>&1; echo "tests rc=$?" >> $S
O=gpurun_out/r27_e2e.jsonl; : > $O
for c in 0 2; do
  timeout 300 python bench.py --only-headline --no-cpu --no-streaming --steps 100 --e2e-steps 300 --e2e-chunks $c 2>> gpurun_out/r27_err.log | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
e = d['e2e']
print(json.dumps({'chunks': $c, 'us_compact_host_policy': 65536e6 / e['value'], 'us_compact_dev_policy': 65536e6 / e['value_with_device_policy_between_calls'], 'us_full_host_policy': 65536e6 / e['value_full_result_format'], 'us_e2e_obs': 65536e6 / d['e2e_obs']['value']}))" >> $O
done
for wl in cfg1_shortest_path; do for c in 0 2; do
  timeout 300 python bench.py --workload $wl --only-headline --no-cpu --no-streaming --steps 100 --e2e-steps 300 --e2e-chunks $c 2>> gpurun_out/r27_err.log | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
e = d['e2e']
print(json.dumps({'wl': '$wl', 'chunks': $c, 'us_compact': 65536e6 / e['value']}))" >> $O
done; done
