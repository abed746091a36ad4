"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python profiles/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphenvs_b200 import BatchedGraphEnv

CFG = [("ShortestPath-v0", 10, 20, {}), ("ShortestPath-v0", 100, 300, {}), ("LongestPath-v0", 50, 200, {"parenting": 2}),
       ("LongestPath-v0", 100, 300, {"parenting": 2}), ("SteinerTree-v0", 40, 100, {"n_dests": 39, "is_eval_env": True}),
       ("TSP-v0", 30, 90, {"parenting": 2}), ("TSP-v0", 70, 300, {"parenting": 2}), ("MaxIndependentSet-v0", 40, 100, {}),
       ("MaxIndependentSet-v0", 100, 300, {}), ("DensestSubgraph-v0", 40, 100, {"parenting": 1}),
       ("DensestSubgraph-v0", 100, 300, {"parenting": 1}), ("MulticastRouting-v0", 60, 200, {"parenting": 4, "n_dests": 3}),
       ("MulticastRouting-v0", 60, 200, {"parenting": 2, "n_dests": 3}), ("DistributionCenter-v0", 60, 200, {"parenting": 2})]
for force in (False, True):
    for env_id, N, E, kw in CFG:
        B = 70
        env = BatchedGraphEnv(env_id, B, N, E, auto_reset=True, structural_features=True, force_warp=force, **kw)
        env.generate(seed=1)
        env.reset()
        env.enable_env_clock()
        for t in range(12):
            if t % 2:
                env.step_sampled(5, 0)
            else:
                env.sample_actions(5, 0)
                env.step_async(env.actions_dev)
        env.obs_flat(0, 4)
        env.stats()
        torch.cuda.synchronize()
        print("ok", env_id, N, "general" if force else "fast", flush=True)
print("done")
