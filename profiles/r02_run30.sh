#!/bin/bash
# round-2 GPU call 30: obs kernel with a thread per node for large graphs: parity (every oracle rollout ends with an obs comparison; golden obs hashes), e2e_obs at cfg5 / cfg3 / cfg4
cd $GRAFT_REPO_ROOT
S=gpurun_out/r30_status.txt; : > $S
timeout 1200 python -m pytest tests/test_cuda_oracle.py tests/test_cuda_golden.py tests/test_cuda_streams.py -m gpu -q -x > gpurun_out/r30_tests.log 2>&1; echo "tests rc=$?" >> $S
O=gpurun_out/r30_e2e_obs.jsonl; : > $O
for wl in cfg5_distcenter cfg5_multicast cfg3_mst cfg4_tsp_p1 densest; do
  timeout 400 python bench.py --workload $wl --only-headline --no-cpu --no-streaming --steps 40 --e2e-steps 40 2>> gpurun_out/r30_err.log | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(json.dumps({'wl': '$wl', 'e2e': d['e2e']['value'], 'e2e_obs': d['e2e_obs']['value'], 'obs_bytes_per_env': d['e2e_obs']['obs_bytes_written_per_env']}))" >> $O
done
