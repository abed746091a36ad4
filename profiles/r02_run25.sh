#!/bin/bash
# round-2 GPU call 25: compact host-result format (ge_step_host_compact) + PDL between the slice kernels of the host step
cd $GRAFT_REPO_ROOT
S=gpurun_out/r25_status.txt; : > $S
timeout 900 python -m pytest tests/test_cuda_oracle.py -m gpu -q -x -k "pipelined or sliced" > gpurun_out/r25_tests.log 2>&1; echo "tests rc=$?" >> $S
O=gpurun_out/r25_e2e.jsonl; : > $O
for pdl in 1 0; do for c in 2 3 4; do
  GE_PIPE_PDL=$pdl python bench.py --only-headline --no-cpu --no-streaming --no-e2e-obs --steps 100 --e2e-steps 300 --e2e-chunks $c 2>> gpurun_out/r25_err.log | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
e = d['e2e']
print(json.dumps({'pipe_pdl': $pdl, 'chunks': $c, 'us_compact_host_policy': 65536e6 / e['value'], 'us_compact_dev_policy': 65536e6 / e['value_with_device_policy_between_calls'], 'us_full_host_policy': 65536e6 / e['value_full_result_format'], 'd2h': e['d2h_bytes_per_step']}))" >> $O
done; done
