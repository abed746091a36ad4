#!/bin/bash
# round-2 GPU call 16: end-to-end host step, write-back kernels vs step kernels storing straight into the pinned host arrays
cd $GRAFT_REPO_ROOT
O=gpurun_out/r16_e2e.jsonl; : > $O
for direct in 0 1; do for c in 1 2 4; do
  GE_PIPE_DIRECT=$direct python bench.py --only-headline --no-cpu --no-streaming --no-e2e-obs --steps 100 --e2e-steps 300 --e2e-chunks $c 2>> gpurun_out/r16_err.log | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(json.dumps({'direct': $direct, 'chunks': $c, 'e2e': d['e2e']['value'], 'e2e_dev_policy': d['e2e']['value_with_device_policy_between_calls'], 'us_per_step': 65536e6 / d['e2e']['value'], 'us_dev_policy': 65536e6 / d['e2e']['value_with_device_policy_between_calls']}))" >> $O
done; done
GE_PIPE_DIRECT=1 timeout 600 python -m pytest tests/test_cuda_oracle.py -x -q -m gpu -k "pipelined" > gpurun_out/r16_tests.log 2>&1
