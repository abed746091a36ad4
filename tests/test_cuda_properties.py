"""GPU: size-independent properties of the transition function, checked exhaustively over the action space
on small instances ("mask <=> step does not flag invalid", "done => frozen").
Every replica i of an instance tries action i from the same state; then all replicas follow one trunk action."""
import random

import numpy as np
import pytest
import torch

from graphenvs_b200 import BatchedGraphEnv
from graphenvs_b200.instances import generate_instance

pytestmark = pytest.mark.gpu

CFG = [
    ("ShortestPath-v0", 12, 24, {}), ("LongestPath-v0", 12, 24, {"parenting": 1}), ("LongestPath-v0", 12, 24, {"parenting": 2}),
    ("SteinerTree-v0", 10, 18, {"n_dests": 4}), ("TSP-v0", 10, 24, {"parenting": 1}), ("TSP-v0", 10, 24, {"parenting": 2}),
    ("MaxIndependentSet-v0", 12, 24, {}), ("DensestSubgraph-v0", 14, 30, {"parenting": 1}),
    ("MulticastRouting-v0", 10, 18, {"parenting": 4, "n_dests": 3}), ("MulticastRouting-v0", 10, 18, {"parenting": 2, "n_dests": 3}),
    ("DistributionCenter-v0", 12, 24, {"parenting": 2}), ("PerishableProductDelivery-v0", 12, 24, {"n_products": 3, "parenting": 1}),
    ("ShortestPath-v0", 70, 160, {}), ("TSP-v0", 70, 200, {"parenting": 2}), ("MaxIndependentSet-v0", 70, 160, {}),
]
STATE = ("node_bits", "node_bits2", "edge_bits", "dist32", "bestkey", "head", "cost", "counters", "done", "mask_bits", "mask_bytes", "mask_cnt")


@pytest.mark.parametrize("force_warp", [False, True], ids=["fast", "general"])
@pytest.mark.parametrize("cfg", CFG, ids=["%s-N%d-%s" % (c[0][:-3], c[1], "".join("%s%s" % (k[0], v) for k, v in c[3].items())) for c in CFG])
def test_mask_is_exactly_the_set_of_accepted_actions(cfg, force_warp):
    env_id, N, E, kw = cfg
    R = 3
    probe = BatchedGraphEnv(env_id, 1, N, E, **kw)
    A = probe.desc.A
    env = BatchedGraphEnv(env_id, R * A, N, E, auto_reset=False, force_warp=force_warp, **kw)
    inst = []
    for r in range(R):
        random.seed(300 + r); np.random.seed(300 + r)
        ins = generate_instance(env_id, env.params)
        inst += [ins] * A
    env.load_instances(inst)
    env.reset()
    every = torch.arange(A, dtype=torch.int32, device="cuda").repeat(R)
    rng = np.random.default_rng(0)
    for t in range(3 * N):
        before = {k: env.t[k].clone() for k in STATE if k in env.t}
        mask = env.mask.cpu().numpy()
        done_before = env.t["done"].cpu().numpy().astype(bool)
        heads = env.t["head"].cpu().numpy()
        _, _, info = env.step(every)
        status = info["status"].cpu().numpy()
        for i in range(R * A):
            a = i % A
            if done_before[i]:
                assert status[i] == 2, "a finished env must refuse to step"
            elif env_id == "TSP-v0" and a == 0 and heads[i] == 0:
                assert status[i] == 0                      # tsp.py:203-211 precedes the mask assert
            else:
                assert (status[i] == 0) == bool(mask[i, a]), (env_id, t, i, a, status[i], mask[i, a])
        rejected = torch.from_numpy(status != 0).cuda()
        for k, v in before.items():
            assert torch.equal(env.t[k][rejected], v[rejected]), "a rejected action must leave %s untouched" % k
        for k, v in before.items():                         # roll back, then advance every replica along one trunk
            env.t[k].copy_(v)
        trunk = np.zeros(R * A, dtype=np.int32)
        for r in range(R):
            valid = np.flatnonzero(mask[r * A])
            trunk[r * A:(r + 1) * A] = rng.choice(valid) if (valid.size and not done_before[r * A]) else 0
        env.step(torch.from_numpy(trunk).cuda())
        if env.t["done"].all():
            break
    if env_id != "PerishableProductDelivery-v0":          # (its random walks rarely deliver everything in 3N moves)
        assert env.t["done"].any()
