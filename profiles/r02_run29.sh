#!/bin/bash
# round-2 GPU call 29: progress counters in the DistributionCenter kernel: parity, e2e chunks 0 vs 2 at config 5
cd $GRAFT_REPO_ROOT
S=gpurun_out/r29_status.txt; : > $S
timeout 600 python -m pytest tests/test_cuda_streams.py tests/test_cuda_oracle.py -m gpu -q -x -k "streamed or pipelined or Distribution" > gpurun_out/r29_tests.log 2>&1; echo "tests rc=$?" >> $S
O=gpurun_out/r29_e2e_dc.jsonl; : > $O
for c in 0 2; do
  timeout 400 python bench.py --workload cfg5_distcenter --only-headline --no-cpu --no-streaming --steps 40 --e2e-steps 60 --e2e-chunks $c 2>> gpurun_out/r29_err.log | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
e = d['e2e']
print(json.dumps({'wl': 'cfg5_distcenter', 'envs': 131072, 'chunks': $c, 'us_per_host_step': 131072e6 / e['value'], 'e2e_env_steps_per_s': e['value'], 'us_e2e_obs': 131072e6 / d['e2e_obs']['value'], 'device_us_isolated': 1e3 * d['isolated']['ms_per_step'], 'd2h_bytes': e['d2h_bytes_per_step']}))" >> $O
done
