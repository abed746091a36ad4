"""One ge_features launch per graph size, for ncu:  python profiles/feature_profile.py [workload] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graphenvs_b200 import BatchedGraphEnv
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5_multicast"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
env_id, N, E, kw, _, _, _ = bench.WORKLOADS[wl]
env = BatchedGraphEnv(env_id, B, N, E, structural_features=True, **kw)
env.generate(seed=3)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record(); env.compute_features(); b.record()
torch.cuda.synchronize()
print(wl, B, "us/env", 1e3 * a.elapsed_time(b) / B)
