"""CPU: the seeded host-side instance generator (graphenvs_b200/instances.py) regenerates the
reference's reset(seed) instances bit-for-bit (graph, edge order, weights, terminals, costs)."""
import random

import numpy as np
import pytest

import golden_util as gu
from graphenvs_b200.instances import generate_instance
from graphenvs_b200.spec import check_ctor_args

CASES = list(gu.all_cases())


@pytest.mark.parametrize("case", CASES, ids=[gu.case_id(m) for m, _ in CASES])
def test_regenerates_reference_instance(case):
    m, r = case
    env_id = m["env_id"]
    kw = dict(m["kwargs"])
    p = check_ctor_args(env_id, kw.pop("n_nodes"), kw.pop("n_edges"), kw)
    random.seed(m["seed"])
    np.random.seed(m["seed"])
    ins = generate_instance(env_id, p)
    np.testing.assert_array_equal(ins.links, r["edge_links"])
    if env_id not in ("MaxIndependentSet-v0", "DensestSubgraph-v0"):
        np.testing.assert_array_equal(ins.w64, r["w64"])
    nodes0 = r["nodes0"]
    if env_id in ("ShortestPath-v0", "LongestPath-v0"):
        assert (ins.src, ins.dest) == (m["src"], m["dest"])
    elif env_id == "SteinerTree-v0":
        assert ins.src == m["src"]
        np.testing.assert_array_equal(ins.dests, r["dests"])
    elif env_id == "MulticastRouting-v0":
        np.testing.assert_array_equal(ins.dests, r["dests"])
        assert 0.0 <= ins.u01 < 1.0
    elif env_id == "DistributionCenter-v0":
        np.testing.assert_array_equal(np.sort(ins.dests), r["targets"])
        np.testing.assert_array_equal(ins.node_cost.astype(np.float32), nodes0[:, 0])
    elif env_id == "MaxIndependentSet-v0":
        np.testing.assert_array_equal(ins.node_cost.astype(np.float32), nodes0[:, 0])
    elif env_id == "TSP-v0" and m["kwargs"].get("spatial"):
        np.testing.assert_array_equal(ins.node_xy.astype(np.float32), nodes0[:, 2:4])
    # the action stream that follows reset() must also line up: same next numpy draw
    if m["policy"] == "random" and len(r["actions"]):
        valid = r["mask0"].nonzero()[0]
        assert int(np.random.choice(valid)) == int(r["actions"][0])


def test_ctor_rules():
    with pytest.raises(AssertionError):
        check_ctor_args("LongestPath-v0", 10, 20, {})            # default parenting=-1 rejected
    with pytest.raises(AssertionError):
        check_ctor_args("TSP-v0", 10, 20, {})
    with pytest.raises(AssertionError):
        check_ctor_args("ShortestPath-v0", 10, 20, {"parenting": 1})
    with pytest.raises(ValueError):
        check_ctor_args("MulticastRouting-v0", 10, 20, {"parenting": 7})
    with pytest.raises(TypeError):
        check_ctor_args("MaxIndependentSet-v0", 10, 20, {"parenting": 1})
    assert check_ctor_args("LongestPath-v0", 50, -1, {"parenting": 2})["n_edges"] == 367
    assert check_ctor_args("DensestSubgraph-v0", 10, 20, {"parenting": 1})["n_choices"] == 3
    assert check_ctor_args("DistributionCenter-v0", 500, 4000, {})["target_count"] == 100


def _heur_golden():
    import json
    import os
    return json.load(open(os.path.join(gu.GOLDEN_DIR, "heuristics.json")))


HG = [h for h in _heur_golden() if h["env_id"] == "MulticastRouting-v0"]


@pytest.mark.parametrize("h", HG, ids=["MC-N%d-s%d" % (h["kwargs"]["n_nodes"], h["seed"]) for h in HG])
def test_multicast_union_of_paths_heuristic_is_bit_identical(h):
    """multicast_routing.py:107-115 depends on networkx's Dijkstra tie order; the host restatement
    (oracle/host_heuristics.multicast_union_of_paths, the checker of the device kernel) must give the reference's
    float64 value exactly."""
    from oracle import host_heuristics as hh
    kw = dict(h["kwargs"])
    p = check_ctor_args(h["env_id"], kw.pop("n_nodes"), kw.pop("n_edges"), kw)
    random.seed(h["seed"])
    np.random.seed(h["seed"])
    ins = generate_instance(h["env_id"], p)
    assert hh.reference_heuristic(h["env_id"], p, ins) == h["heuristic"]


HN = [h for h in _heur_golden() if h["env_id"] in ("TSP-v0", "MaxIndependentSet-v0")
      or (h["env_id"] == "SteinerTree-v0" and 1 < h["kwargs"]["n_dests"] < h["kwargs"]["n_nodes"] - 1)]


@pytest.mark.parametrize("h", HN, ids=["%s-N%d-s%d" % (h["env_id"][:-3], h["kwargs"]["n_nodes"], h["seed"]) for h in HN])
def test_networkx_defined_heuristics_match_reference(h):
    """Kou / Christofides / Ramsey values are defined by networkx's iteration order; the oracle-side delegate
    (oracle/host_heuristics.py) rebuilds the reference's nx.Graph insertion order from the host generator's edge order and
    must return the reference's value (1e-9: only the float summation order of the final edge list may differ).  This pins
    the host generator's edge order; the product reports labelled device alternatives for these three."""
    pytest.importorskip("networkx")
    import warnings
    from oracle import host_heuristics as hh
    warnings.filterwarnings("ignore")
    kw = dict(h["kwargs"])
    p = check_ctor_args(h["env_id"], kw.pop("n_nodes"), kw.pop("n_edges"), kw)
    random.seed(h["seed"])
    np.random.seed(h["seed"])
    ins = generate_instance(h["env_id"], p)
    assert hh.reference_heuristic(h["env_id"], p, ins) == pytest.approx(h["heuristic"], rel=1e-9, abs=1e-12)


def test_gnm_generator_matches_networkx_draw_for_draw():
    """gnm_adjacency == nx.gnm_random_graph under the same `random` seed: same edges, same adjacency insertion
    order (the edge-action index of SteinerTree / Multicast is a position in that order), same RNG consumption."""
    nx = pytest.importorskip("networkx")
    from graphenvs_b200.instances import gnm_adjacency
    rng = np.random.default_rng(5)
    for _ in range(60):
        n = int(rng.integers(2, 60))
        m = int(rng.integers(0, n * (n - 1) // 2 + 3))
        seed = int(rng.integers(0, 10_000))
        random.seed(seed)
        G = nx.gnm_random_graph(n, m)
        after_nx = random.random()
        random.seed(seed)
        order = []
        adj = gnm_adjacency(n, m, order)
        after_mine = random.random()
        assert after_nx == after_mine, "different number of random draws"
        assert [list(G.adj[u]) for u in range(n)] == [list(a) for a in adj]
        H = nx.Graph()
        H.add_nodes_from(range(n))
        H.add_edges_from(order)
        assert [list(H.adj[u]) for u in range(n)] == [list(G.adj[u]) for u in range(n)], "edge order must rebuild G exactly"
