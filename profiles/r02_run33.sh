#!/bin/bash
# round-2 GPU call 33: 128-bit copy-out in the thread-per-node observation kernel: parity (obs compared at the end of every oracle rollout), timing
cd $GRAFT_REPO_ROOT
S=gpurun_out/r33_status.txt; : > $S
timeout 900 python -m pytest tests/test_cuda_oracle.py -m gpu -q -x -k "random_rollout and (N300 or N500 or N600 or N1100 or N200 or N120 or N130) or obs" > gpurun_out/r33_tests.log 2>&1; echo "tests rc=$?" >> $S
python - > gpurun_out/r33_obs_timing.jsonl 2>> gpurun_out/r33_err.log <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
import bench
from graphenvs_b200 import BatchedGraphEnv
for wl in ("cfg5_distcenter", "cfg5_multicast", "densest"):
    env_id, N, E, kw, B, _, _ = bench.WORKLOADS[wl]
    e = BatchedGraphEnv(env_id, B, N, E, auto_reset=True, **kw)
    e.generate(seed=1); e.reset()
    x = torch.empty((B, N, e.F), dtype=torch.float32, device="cuda")
    for _ in range(3): e.obs_nodes(out=x)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): e.obs_nodes(out=x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(json.dumps({"wl": wl, "obs_nodes_ms": ms, "bytes": x.numel() * 4, "GBps": x.numel() * 4 / ms / 1e6}))
    del e, x; torch.cuda.empty_cache()
PY
echo "timing rc=$?" >> $S
