"""GPU: the device instance generator (ge_generate) produces valid instances of the reference's
distribution (connected simple G(n,m), symmetric weights in {0.3..0.9}, distinct terminals; TSP's
extra rejections tsp.py:60-71; DensestSubgraph's isolated stop node densest_subgraph.py:59-65) and
is a pure function of (seed, global env id): a rank-sliced batch equals the single-GPU batch."""
import numpy as np
import pytest
import torch

from graphenvs_b200 import BatchedGraphEnv

pytestmark = pytest.mark.gpu

CFG = [
    ("ShortestPath-v0", 10, 20, {}),
    ("LongestPath-v0", 50, 200, {"parenting": 2}),
    ("SteinerTree-v0", 100, 500, {"n_dests": 99}),
    ("SteinerTree-v0", 40, 100, {"n_dests": 4}),
    ("TSP-v0", 30, 80, {"parenting": 2}),
    ("TSP-v0", 20, 190, {"parenting": 1}),
    ("MaxIndependentSet-v0", 70, 300, {}),
    ("DensestSubgraph-v0", 60, 200, {"parenting": 1}),
    ("MulticastRouting-v0", 120, 600, {"n_dests": 5, "parenting": 4}),
    ("DistributionCenter-v0", 100, 400, {"parenting": 2}),
]


def _connected(adj, nodes):
    nodes = list(nodes)
    seen = {nodes[0]}
    stack = [nodes[0]]
    allowed = set(nodes)
    while stack:
        u = stack.pop()
        for v in adj[u]:
            if v in allowed and v not in seen:
                seen.add(v); stack.append(v)
    return len(seen) == len(nodes)


@pytest.mark.parametrize("cfg", CFG, ids=["%s-N%d-E%d" % (c[0][:-3], c[1], c[2]) for c in CFG])
def test_generated_instances_are_valid(cfg):
    env_id, N, E, kw = cfg
    B = 48
    env = BatchedGraphEnv(env_id, B, N, E, **kw)
    env.generate(seed=5)
    torch.cuda.synchronize()
    inst = env.export_instances()
    n_graph = N - 1 if env_id == "DensestSubgraph-v0" else N
    for ins in inst:
        links = ins.links
        assert links.shape == (2 * E, 2)
        assert np.all(links[1:, 0] >= links[:-1, 0])
        assert np.all(links[:, 0] != links[:, 1])
        pairs = set(map(tuple, links.tolist()))
        assert len(pairs) == 2 * E, "duplicate directed edge"
        wmap = {tuple(l): w for l, w in zip(links.tolist(), ins.w64)}
        for (u, v), w in wmap.items():
            assert (v, u) in wmap and wmap[(v, u)] == w, "asymmetric edge / weight"
        if env_id in ("MaxIndependentSet-v0", "DensestSubgraph-v0"):
            assert np.all(ins.w64 == 1.0)
        else:
            assert set(np.round(ins.w64 * 10).astype(int)) <= set(range(3, 10))
            assert np.all(ins.w64 == np.round(ins.w64 * 10) / 10.0), "weights must be the fp64 values k/10"
        adj = [[] for _ in range(N)]
        for u, v in links:
            adj[u].append(v)
        assert _connected(adj, range(n_graph))
        if env_id == "DensestSubgraph-v0":
            assert len(adj[N - 1]) == 0
        if env_id == "TSP-v0":
            assert min(len(a) for a in adj) >= 2 or E >= N * (N - 1) // 2
            assert _connected(adj, range(1, N))
        if env_id in ("ShortestPath-v0", "LongestPath-v0"):
            assert ins.src != ins.dest and 0 <= ins.src < N and 0 <= ins.dest < N
        if env_id == "SteinerTree-v0":
            assert len(ins.dests) == kw["n_dests"] and ins.src not in ins.dests
        if env_id == "MulticastRouting-v0":
            assert len(ins.dests) == kw["n_dests"] and 0 not in ins.dests and ins.max_distance > 0
        if env_id == "DistributionCenter-v0":
            assert len(set(ins.dests.tolist())) == env.desc.n_targets
            assert set(ins.node_cost.tolist()) <= {1.0, 2.0, 3.0}
        if env_id == "MaxIndependentSet-v0":
            assert set(np.round(ins.node_cost * 10).astype(int)) <= set(range(3, 10))
    # distribution sanity: weights roughly uniform over the 7 values (B*E draws)
    if env_id in ("LongestPath-v0", "SteinerTree-v0") and N >= 50:
        allw = np.concatenate([np.round(i.w64 * 10).astype(int) for i in inst])
        freq = np.bincount(allw, minlength=10)[3:10] / allw.size
        assert np.all(np.abs(freq - 1 / 7) < 0.02)


def test_rank_sliced_generation_equals_single_batch():
    full = BatchedGraphEnv("LongestPath-v0", 64, 50, 200, parenting=2, auto_reset=True)
    full.generate(seed=9)
    full.reset()
    halves = []
    for r in range(2):
        h = BatchedGraphEnv("LongestPath-v0", 32, 50, 200, parenting=2, auto_reset=True, env_id0=32 * r)
        h.generate(seed=9)
        h.reset()
        halves.append(h)
    for t in range(30):
        for e in [full] + halves:
            e.sample_actions(77, t)
            e.step_async(e.actions_dev)
    torch.cuda.synchronize()
    for name in ("row_ptr", "col", "w64", "src", "dest", "node_bits", "mask_bits", "head", "cost", "acc", "traj"):
        a = full.t[name]
        b = torch.cat([h.t[name] for h in halves], dim=1 if name == "acc" else 0)
        assert torch.equal(a, b), name
