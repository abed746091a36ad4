"""GPU: reset-time derived data computed by the CUDA engine against the reference's recorded
values (tests/golden): structural features (feature_extraction.py:6-37), tie-independent eval
heuristics (shortest_path.py:90, longest_path.py:105, steiner_tree.py:79,81), DistributionCenter
in-range tables (distribution_center.py:113-116), Multicast max_distance (multicast_routing.py:98-103)."""
import numpy as np
import pytest
import torch

import cuda_util as cu
import golden_util as gu
from graphenvs_b200 import BatchedGraphEnv

pytestmark = pytest.mark.gpu
GROUPS = cu.grouped_cases()


def _bare_instance(m, r):
    ins = cu.instance_from_case(m, r)
    ins.features = None
    ins.heuristic = None
    return ins


@pytest.mark.parametrize("key", list(GROUPS), ids=["-".join(str(x) for x in k) for k in GROUPS])
def test_features_and_heuristics(key):
    cases = GROUPS[key]
    m0 = cases[0][0]
    kw = dict(m0["kwargs"])
    n_nodes, n_edges = kw.pop("n_nodes"), kw.pop("n_edges")
    kw["is_eval_env"] = True
    env = BatchedGraphEnv(m0["env_id"], len(cases), n_nodes, n_edges, structural_features=True, **kw)
    env.load_instances([_bare_instance(m, r) for m, r in cases])
    torch.cuda.synchronize()
    feats = env.t["features"].cpu().numpy()
    heur = env.t["heuristic"].cpu().numpy()
    nd = gu.DYN_COLS[m0["env_id"]]
    for b, (m, r) in enumerate(cases):
        ref32 = r["nodes0"][:, nd:]
        # tolerance of the test: 1e-5 relative in fp32 (north_star); degree column exact
        np.testing.assert_array_equal(feats[b][:, 0], ref32[:, 0])
        np.testing.assert_allclose(feats[b], ref32, rtol=1e-5, atol=1e-8, err_msg="features env %d" % b)
        np.testing.assert_allclose(feats[b], r["features64"].astype(np.float32), rtol=1e-5, atol=1e-8)
        if env.spec.heuristic_on_device(env.params) and m["kwargs"].get("is_eval_env"):
            assert heur[b] == pytest.approx(m["heuristic"], rel=1e-9), "heuristic env %d" % b
    if m0["env_id"] == "DistributionCenter-v0" and env.desc.parenting == 2:
        tab = env.t["in_range"].cpu().numpy().view(np.uint32)
        for b, (m, r) in enumerate(cases):
            bits = np.unpackbits(tab[b].view(np.uint8).reshape(tab.shape[1], -1), axis=1, bitorder="little")[:, :m["N"]]
            # engine rows follow `targets` order = sorted target ids (golden_util) = r["in_range"] rows
            np.testing.assert_array_equal(bits, r["in_range"], err_msg="in_range env %d" % b)


def test_features_large_vs_oracle():
    """N beyond the fixtures (multi-word bitsets, deeper BFS): CUDA features vs the C oracle's fp64 values."""
    import random
    from graphenvs_b200.instances import generate_instance
    for env_id, N, E, kw in [("ShortestPath-v0", 150, 400, {}), ("TSP-v0", 60, 400, {"parenting": 1}),
                             ("MaxIndependentSet-v0", 200, 5970, {}), ("DensestSubgraph-v0", 90, 300, {"parenting": 1}),
                             ("TSP-v0", 200, 19900, {"parenting": 1}),               # config 4: complete graph, weighted pagerank
                             ("MulticastRouting-v0", 500, 4000, {"parenting": 4}),   # config 5 graph size (16 set words)
                             ("ShortestPath-v0", 1000, 2500, {}),                     # 32 set words, long BFS levels
                             ("ShortestPath-v0", 1100, 3000, {}),                     # N > 1024: the warp-per-env kernel
                             ("ShortestPath-v0", 9, 12, {}), ("DensestSubgraph-v0", 40, 45, {"parenting": 0})]:   # near-trees: deep levels
        B = 6 if N < 500 else 3
        env = BatchedGraphEnv(env_id, B, N, E, structural_features=True, **kw)
        inst = []
        for b in range(B):
            random.seed(50 + b); np.random.seed(50 + b)
            inst.append(generate_instance(env_id, env.params))
        env.load_instances(inst)
        torch.cuda.synchronize()
        feats = env.t["features"].cpu().numpy()
        for b, ins in enumerate(inst):
            oe = cu.oracle_from_instance(env_id, ins, env.params)
            f = oe.features64(weighted_pr=(env_id == "TSP-v0")).astype(np.float32)
            np.testing.assert_allclose(feats[b], f, rtol=1e-5, atol=1e-8, err_msg="%s env %d" % (env_id, b))


def test_device_heuristics_match_reference_values_beyond_fixture_sizes():
    """tests/golden/heuristics.json (oracle/gen_heuristic_golden.py): Dijkstra / MST values of the reference's
    reset(seed) for N = 60-80, recomputed on the device from the host-regenerated instances."""
    import json, os, random
    from graphenvs_b200.instances import generate_instance
    hs = json.load(open(os.path.join(gu.GOLDEN_DIR, "heuristics.json")))
    groups = {}
    for h in hs:
        kwh = h["kwargs"]
        if h["env_id"] == "ShortestPath-v0" or (h["env_id"] == "SteinerTree-v0" and kwh["n_dests"] in (1, kwh["n_nodes"] - 1)):
            groups.setdefault((h["env_id"], json.dumps(h["kwargs"], sort_keys=True)), []).append(h)
    assert groups
    for (env_id, kws), lst in groups.items():
        kw = json.loads(kws)
        n_nodes, n_edges = kw.pop("n_nodes"), kw.pop("n_edges")
        env = BatchedGraphEnv(env_id, len(lst), n_nodes, n_edges, **kw)
        inst = []
        for h in lst:
            random.seed(h["seed"]); np.random.seed(h["seed"])
            inst.append(generate_instance(env_id, env.params))
        env.load_instances(inst)
        torch.cuda.synchronize()
        got = env.t["heuristic"].cpu().numpy()
        for b, h in enumerate(lst):
            assert got[b] == pytest.approx(h["heuristic"], rel=1e-9), (env_id, kw, h["seed"])


def test_features_cta_kernel_equals_warp_kernel():
    """Round-2 CTA-per-env features kernel vs the round-1 warp-per-env kernel (force_warp): same float32 columns up to
    fp64 reassociation."""
    import random
    from graphenvs_b200.instances import generate_instance
    for env_id, N, E, kw in [("LongestPath-v0", 50, 200, {"parenting": 2}), ("SteinerTree-v0", 100, 500, {"n_dests": 99}),
                             ("DistributionCenter-v0", 300, 1200, {"parenting": 2})]:
        inst = []
        for b in range(16):
            random.seed(90 + b); np.random.seed(90 + b)
            inst.append(generate_instance(env_id, BatchedGraphEnv(env_id, 1, N, E, **kw).params))
        got = []
        for fw in (False, True):
            env = BatchedGraphEnv(env_id, 16, N, E, structural_features=True, force_warp=fw, **kw)
            env.load_instances(inst)
            torch.cuda.synchronize()
            got.append(env.t["features"].cpu().numpy())
        np.testing.assert_allclose(got[0], got[1], rtol=1e-6, atol=1e-9, err_msg=env_id)
