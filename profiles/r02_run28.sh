#!/bin/bash
# round-2 GPU call 28: streamed write-back where it is meant for -- a multi-wave batch (cfg2 shape, 1,048,576 envs): chunks 0 / 2 / 4
cd $GRAFT_REPO_ROOT
O=gpurun_out/r28_e2e_1m.jsonl; : > $O
for c in 0 2 4; do
  timeout 400 python bench.py --only-headline --no-cpu --no-streaming --no-e2e-obs --envs 1048576 --steps 40 --e2e-steps 60 --e2e-chunks $c --e2e-device-policy 2>> gpurun_out/r28_err.log | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
e = d['e2e']
print(json.dumps({'envs': 1048576, 'chunks': $c, 'us_per_host_step': 1048576e6 / e['value'], 'e2e_env_steps_per_s': e['value'], 'device_us_isolated': 1e3 * d['isolated']['ms_per_step'], 'd2h_bytes': e['d2h_bytes_per_step']}))" >> $O
done
