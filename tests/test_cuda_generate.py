"""GPU: the device instance generator (ge_generate) produces valid instances of the reference's
distribution (connected simple G(n,m), symmetric weights in {0.3..0.9}, distinct terminals; TSP's
extra rejections tsp.py:60-71; DensestSubgraph's isolated stop node densest_subgraph.py:59-65) and
is a pure function of (seed, global env id): a rank-sliced batch equals the single-GPU batch."""
import numpy as np
import pytest
import torch

import cuda_util as cu
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
    ("PerishableProductDelivery-v0", 50, 200, {"n_products": 3, "parenting": 1}),
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
        if env_id == "PerishableProductDelivery-v0":          # perishable_product_delivery.py:96-114
            P = kw["n_products"]
            assert len(set(ins.dests.tolist())) == 2 * P, "pickups and dropoffs are distinct nodes"
            assert env.params["dt_mn"] - 1e-6 <= ins.max_distance <= env.params["dt_mx"] + 1e-6
            oe = cu.oracle_from_instance(env_id, ins, env.params)
            for i in range(P):
                assert oe.sssp(int(ins.dests[i]))[int(ins.dests[P + i])] < ins.max_distance + 1e-5, "dropoff within the delivery time of its pickup"
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


def test_rank_sliced_multicast_generation_equals_single_batch():
    """the Multicast max_distance uniform must be a function of the GLOBAL env id, not of the rank's batch."""
    kw = dict(n_dests=4, parenting=4)
    full = BatchedGraphEnv("MulticastRouting-v0", 96, 60, 200, auto_reset=True, **kw)
    full.generate(seed=9)
    full.reset()
    parts = []
    for r in range(3):
        h = BatchedGraphEnv("MulticastRouting-v0", 32, 60, 200, auto_reset=True, env_id0=32 * r, **kw)
        h.generate(seed=9)
        h.reset()
        parts.append(h)
    for t in range(30):
        for e in [full] + parts:
            e.sample_actions(77, t)
            e.step_async(e.actions_dev)
    torch.cuda.synchronize()
    md = full.t["max_dist32"]
    assert len(torch.unique(md)) > 48, "per-env draws"
    for name in ("row_ptr", "col", "w32", "max_dist32", "target_bits", "node_bits", "mask_bits", "dist32", "cost", "acc", "traj"):
        a = full.t[name]
        b = torch.cat([h.t[name] for h in parts], dim=1 if name == "acc" else 0)
        assert torch.equal(a, b), name


def _edge_order_is_consistent(links, N):
    """True iff ONE order of the undirected edges explains every row: row u lists its neighbours in the order the edges
    touching u were inserted (list(G.to_directed().edges) of the reference).  Kahn's algorithm over the constraints
    'edge of row slot i precedes edge of row slot i+1'."""
    from collections import defaultdict, deque
    rows = defaultdict(list)
    for u, v in links.tolist():
        rows[u].append((min(u, v), max(u, v)))
    succ, indeg = defaultdict(list), defaultdict(int)
    nodes = set()
    for u, es in rows.items():
        nodes.update(es)
        for a, b in zip(es[:-1], es[1:]):
            succ[a].append(b)
            indeg[b] += 1
    q = deque(e for e in nodes if indeg[e] == 0)
    seen = 0
    while q:
        e = q.popleft()
        seen += 1
        for f in succ[e]:
            indeg[f] -= 1
            if indeg[f] == 0:
                q.append(f)
    return seen == len(nodes)


def test_rows_are_in_insertion_order_like_the_reference():
    """rows were neighbour-sorted; the reference's are insertion-ordered (decides np.argmin ties under
    Multicast parenting >= 3).  The device rows must be explainable by one global edge order, and must NOT be sorted."""
    import random
    from graphenvs_b200.instances import generate_instance
    env = BatchedGraphEnv("MulticastRouting-v0", 64, 120, 600, n_dests=5, parenting=4)
    env.generate(seed=2)
    torch.cuda.synchronize()
    unsorted_rows = total_rows = 0
    for ins in env.export_instances():
        assert _edge_order_is_consistent(ins.links, 120)
        rp = np.searchsorted(ins.links[:, 0], np.arange(121))
        for u in range(120):
            r = ins.links[rp[u]:rp[u + 1], 1]
            total_rows += 1
            unsorted_rows += int(np.any(r[1:] < r[:-1]))
    assert unsorted_rows > 0.8 * total_rows, "rows look sorted: %d of %d unsorted" % (unsorted_rows, total_rows)
    # the host generator (= the reference's draws) passes the same consistency check, i.e. the check is the right one
    random.seed(1); np.random.seed(1)
    ref = generate_instance("MulticastRouting-v0", env.params)
    assert _edge_order_is_consistent(ref.links, 120)


def test_generated_distribution_matches_host_generator():
    """Distribution parity with the reference's reset(): degree histogram, position of a neighbour id inside its row
    (uniform for insertion order), terminal uniformity -- device generator vs the host generator that replays the
    reference's own draws."""
    import random
    from graphenvs_b200.instances import generate_instance
    env_id, N, E, kw = "SteinerTree-v0", 40, 100, {"n_dests": 4}
    B = 2048
    env = BatchedGraphEnv(env_id, B, N, E, **kw)
    env.generate(seed=17)
    torch.cuda.synchronize()
    rp = env.t["row_ptr"][:, :N + 1].cpu().numpy()
    deg_dev = np.bincount(np.diff(rp, axis=1).ravel(), minlength=24)[:24] / (B * N)
    tb = env.t["target_bits"].cpu().numpy().view(np.uint32)
    tgt = np.unpackbits(tb.view(np.uint8), axis=1, bitorder="little")[:, :N].sum(0) / B
    src = np.bincount(env.t["src"].cpu().numpy(), minlength=N) / B
    deg_host = np.zeros(24)
    H = 300
    for b in range(H):
        random.seed(500 + b); np.random.seed(500 + b)
        ins = generate_instance(env_id, env.params)
        deg_host += np.bincount(np.bincount(ins.links[:, 0], minlength=N), minlength=24)[:24]
    deg_host /= H * N
    assert np.abs(deg_dev - deg_host).max() < 0.02, (deg_dev, deg_host)
    assert np.abs(tgt - kw["n_dests"] / N).max() < 0.03 and np.abs(src - 1 / N).max() < 0.02
    # first slot of a row: under insertion order the smallest neighbour id is first only ~1/deg of the time
    col = env.t["col"].cpu().numpy()
    first_is_min = 0
    rows = 0
    for b in range(256):
        for u in range(N):
            r = col[b, rp[b, u]:rp[b, u + 1]]
            if r.size >= 3:
                rows += 1
                first_is_min += int(r[0] == r.min())
    assert first_is_min / rows < 0.45


def test_sparse_configuration_falls_back_to_connected_by_construction():
    """when 4096 rejection draws find no valid graph (a tree-sparse G(n, n-1) is almost never connected) the
    generator used to emit the last rejected graph.  Now every env is valid and the fallbacks are counted."""
    import warnings
    for env_id, N, E, kw in [("SteinerTree-v0", 60, 59, {"n_dests": 3}), ("TSP-v0", 40, 41, {"parenting": 1})]:
        env = BatchedGraphEnv(env_id, 32, N, E, auto_reset=True, **kw)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            env.generate(seed=3)
        assert env.generate_fallbacks > 0 and any("connected-by-construction" in str(x.message) for x in w)
        for ins in env.export_instances():
            adj = [[] for _ in range(N)]
            for u, v in ins.links:
                adj[u].append(v)
            assert len(set(map(tuple, ins.links.tolist()))) == 2 * E
            assert _connected(adj, range(N))
            if env_id == "TSP-v0":
                assert min(len(a) for a in adj) >= 2 and _connected(adj, range(1, N))
        env.reset()
        for t in range(20):          # the envs step and finish episodes
            env.step_sampled(1, t)
        torch.cuda.synchronize()
        assert int(env.flags[:, 2].max()) == 0
    ok = BatchedGraphEnv("LongestPath-v0", 256, 50, 200, parenting=2)
    ok.generate(seed=3)
    assert ok.generate_fallbacks == 0


def test_spatial_tsp_generation_is_euclidean():
    """spatial=True got zero coordinates and k/10 weights; tsp.py:80-86 draws U(0,10)^2 and uses distances."""
    env = BatchedGraphEnv("TSP-v0", 16, 30, 80, parenting=1, spatial=True)
    env.generate(seed=4)
    torch.cuda.synchronize()
    xy = env.t["node_xy"].cpu().numpy().astype(np.float64)
    assert xy.min() >= 0 and xy.max() <= 10 and xy.std() > 2
    for b, ins in enumerate(env.export_instances()):
        u, v = ins.links[:, 0], ins.links[:, 1]
        dist = np.sqrt(((xy[b, u] - xy[b, v]) ** 2).sum(1))
        np.testing.assert_allclose(ins.w64, dist, rtol=1e-5)     # xy is stored as float32, the weight came from fp64 coordinates


@pytest.mark.parametrize("cfg", [("ShortestPath-v0", 10, 20, {}), ("LongestPath-v0", 50, 200, {"parenting": 2}),
                                 ("SteinerTree-v0", 40, 100, {"n_dests": 4}), ("TSP-v0", 20, 60, {"parenting": 1}),
                                 ("DistributionCenter-v0", 100, 400, {"parenting": 2}), ("MulticastRouting-v0", 60, 200, {"n_dests": 3, "parenting": 4})],
                         ids=lambda c: c[0][:-3])
def test_instance_pool_gives_finished_envs_a_fresh_instance(cfg):
    """Regenerate-on-done: after turn_over() every finished env carries an instance of a pool bank (bit-equal
    arrays), is reset (done = 0, fresh mask), and the other envs are untouched; with background regeneration the banks
    change over time.  The stepped trajectories stay valid (no invalid action statuses)."""
    from graphenvs_b200.pool import InstancePool
    env_id, N, E, kw = cfg
    B = 300
    env = BatchedGraphEnv(env_id, B, N, E, auto_reset=False, **kw)
    env.generate(seed=1)
    env.reset()
    pool = InstancePool(env, banks=3, seed=7, background=False)
    first_col = env.t["col"].clone()
    seen_done = 0
    for t in range(60):
        env.step_sampled(5, t)
        done = env.t["done"].clone().bool()
        before = {k: env.t[k].clone() for k in ("col", "row_ptr", "node_bits", "mask_bits")}
        ep_before = pool.episode.clone()
        pool.turn_over()
        torch.cuda.synchronize()
        assert int(env.flags[:, 2].max()) == 0
        assert not env.t["done"].any(), "every finished env has been reset"
        assert torch.equal(pool.select.bool(), done)
        for k, v in before.items():
            assert torch.equal(env.t[k][~done], v[~done]), "%s of an unfinished env changed" % k
        if done.any():
            seen_done += int(done.sum())
            idx = torch.nonzero(done).flatten()
            bank_of = (ep_before[idx] % pool.n_active).cpu().tolist()
            for i, k in zip(idx.cpu().tolist()[:20], bank_of[:20]):
                bk = pool.banks[pool._host_order[k]]
                assert torch.equal(env.t["col"][i], bk.t["col"][i]) and torch.equal(env.t["row_ptr"][i], bk.t["row_ptr"][i])
            assert torch.equal(pool.episode[idx], ep_before[idx] + 1)
    assert seen_done > B, "episodes must have turned over"
    assert not torch.equal(env.t["col"], first_col)
    # the refilled envs behave like freshly loaded ones: a twin batch loaded with the same instances agrees step for step
    twin = BatchedGraphEnv(env_id, B, N, E, auto_reset=False, **kw)
    twin.load_instances(env.export_instances())
    twin.reset(select=torch.ones(B, dtype=torch.uint8))
    env.reset()
    for t in range(10):
        a = env.sample_actions(9, t).clone()
        env.step_async(a)
        twin.step_async(a)
    torch.cuda.synchronize()
    assert torch.equal(env.t["mask_bits"], twin.t["mask_bits"]) and torch.equal(env.reward, twin.reward)


def test_instance_pool_background_regeneration_swaps_banks():
    from graphenvs_b200.pool import InstancePool
    env = BatchedGraphEnv("ShortestPath-v0", 2048, 10, 20, auto_reset=False)
    env.generate(seed=1)
    env.reset()
    pool = InstancePool(env, banks=3, seed=3, background=True)
    snap = [b.t["col"].clone() for b in pool.banks]
    for t in range(400):
        env.step_sampled(5, t)
        pool.turn_over()
        if t % 50 == 0:
            torch.cuda.synchronize()
    pool.close()
    torch.cuda.synchronize()
    assert pool.regenerated >= 1, "the background stream must have delivered at least one new bank"
    assert sum(int(not torch.equal(s, b.t["col"])) for s, b in zip(snap, pool.banks)) >= 1
    assert int(env.flags[:, 2].max()) == 0
