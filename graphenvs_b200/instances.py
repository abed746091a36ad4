"""Host-side, seeded instance generation for the gymnasium-shaped single-env API.

Reproduces the instance the reference's `reset(seed)` builds, draw for draw, WITHOUT networkx:
the reference reseeds the process-global `random` and `numpy.random` (shortest_path.py:49-52),
builds G(n,m) with two `random.choice(nlist)` per attempt (nx:generators/random_graphs.py:298-310),
then draws weights / terminals from the global numpy stream in the order of SURVEY.md Appendix A.
Like the reference, this module therefore uses (and advances) the GLOBAL generators.

Edge order contract: `links` = list(DiGraph.edges) of the reference = source-sorted rows, each row
in neighbour insertion order.
"""
import random
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class Instance:
    n_nodes: int
    links: np.ndarray                     # int32 [M, 2], reference order
    w64: np.ndarray                       # float64 [M]  edge attribute ('delay' / 'weight'); 1.0 when unused
    src: int = 0
    dest: int = 0
    dests: Optional[np.ndarray] = None    # Steiner/Multicast destinations, DistributionCenter targets
    node_cost: Optional[np.ndarray] = None
    node_xy: Optional[np.ndarray] = None
    max_distance: Optional[float] = None  # Multicast: value of the MAX_DISTANCE column (None -> from u01 on device)
    u01: Optional[float] = None           # Multicast: the reference's np.random.rand() draw
    heuristic: Optional[float] = None     # None -> computed on the device when is_eval_env
    edge_order: Optional[list] = None     # undirected edges in the order gnm_random_graph inserted them
    features: Optional[np.ndarray] = None  # float32 [N, 5]; None -> computed on the device


def _connected(adj, n, skip=-1):
    """BFS connectivity of the graph on nodes {0..n-1} minus `skip`."""
    start = 0 if skip != 0 else 1
    if n - (1 if skip >= 0 else 0) <= 0:
        return False
    seen = [False] * n
    seen[start] = True
    stack = [start]
    cnt = 1
    while stack:
        u = stack.pop()
        for v in adj[u]:
            if v != skip and not seen[v]:
                seen[v] = True
                cnt += 1
                stack.append(v)
    return cnt == n - (1 if skip >= 0 else 0)


def gnm_adjacency(n, m, edge_order=None):
    """nx.gnm_random_graph(n, m) with seed=None: adjacency lists in insertion order.  `edge_order`, when given,
    receives the (u, v) pairs in the order networkx inserted them (enough to rebuild an identical nx.Graph)."""
    adj = [dict() for _ in range(n)]
    if n == 1:
        return adj
    if m >= n * (n - 1) / 2.0:  # complete_graph: itertools.combinations order, no randomness consumed
        for u in range(n):
            for v in range(u + 1, n):
                adj[u][v] = None
                adj[v][u] = None
                if edge_order is not None:
                    edge_order.append((u, v))
        return adj
    nlist = list(range(n))
    cnt = 0
    choice = random.choice
    while cnt < m:
        u = choice(nlist)
        v = choice(nlist)
        if u == v or v in adj[u]:
            continue
        adj[u][v] = None
        adj[v][u] = None
        if edge_order is not None:
            edge_order.append((u, v))
        cnt += 1
    return adj


def _undirected_edges(adj):
    """Order of G.edges on the undirected graph: (u, v) yielded from the endpoint seen first."""
    seen = set()
    for u in range(len(adj)):
        for v in adj[u]:
            if v not in seen:
                yield u, v
        seen.add(u)


def _perishable_instance(p):
    """PerishableProductDelivery reset() (perishable_product_delivery.py:72-117), draw for draw: the numpy draws (weight
    matrix, delivery time, pickups / dropoffs) happen INSIDE the retry loop, the pickup / dropoff lists are NOT cleared
    between attempts (:78-79 sit before the loop), and a dropoff is drawn from the keys of nx.floyd_warshall's dict of the
    pickup in ITS insertion order: the node itself, the far ends of its incident edges in G.edges order, then the
    remaining nodes ascending (nx:algorithms/shortest_paths/dense.py).  Distances are the literal Floyd-Warshall values."""
    N, E, P = p["n_nodes"], p["n_edges"], p["n_products"]
    weighted = p.get("weighted", True)
    rnd = np.random
    pickups, dropoffs = [-1] * P, [-1] * P
    while True:
        edge_order = []
        adj = gnm_adjacency(N, E, edge_order)
        if not _connected(adj, N):
            continue
        delay = (rnd.randint(3, 10, size=(N, N)) if weighted else rnd.randint(10, 11, size=(N, N))) / 10.0
        und = list(_undirected_edges(adj))                          # G.edges order, u < v
        delivery_time = rnd.rand() * (p["dt_mx"] - p["dt_mn"]) + p["dt_mn"]
        dist = np.full((N, N), np.inf)
        np.fill_diagonal(dist, 0.0)
        order = [[u] for u in range(N)]                            # key order of dist[u]
        for u, v in und:
            w = delay[u, v]
            dist[u, v] = min(w, dist[u, v]); dist[v, u] = min(w, dist[v, u])
            order[u].append(v); order[v].append(u)
        for k in range(N):                                          # for w in G: d = dist[u][w] + dist[w][v]
            d = dist[:, k:k + 1] + dist[k:k + 1, :]
            np.minimum(dist, d, out=dist)
        for u in range(N):
            have = set(order[u])
            order[u] += [v for v in range(N) if v not in have]
        ok = True
        for i in range(P):
            pickups[i] = rnd.choice([node for node in range(N) if (node not in pickups) and (node not in dropoffs)])
            close = [node for node in order[pickups[i]] if dist[pickups[i], node] < delivery_time + 1e-6
                     and (node not in pickups) and (node not in dropoffs)]
            if len(close) == 0:
                ok = False
                break
            dropoffs[i] = rnd.choice(close)
        if ok and (-1 not in pickups) and (-1 not in dropoffs):
            break
    links = np.array([(u, v) for u in range(N) for v in adj[u]], dtype=np.int32).reshape(-1, 2)
    lo = np.minimum(links[:, 0], links[:, 1])
    hi = np.maximum(links[:, 0], links[:, 1])
    ins = Instance(n_nodes=N, links=links, w64=delay[lo, hi].astype(np.float64), edge_order=edge_order)
    ins.dests = np.array([int(x) for x in pickups] + [int(x) for x in dropoffs], dtype=np.int32)   # pickups, then dropoffs
    ins.max_distance = float(delivery_time)
    return ins


def generate_instance(env_id, p):
    """Instance of `env_id` with constructor parameters `p` (spec.check_ctor_args), consuming the
    global `random` / `numpy.random` streams exactly like the reference's reset()."""
    if env_id == "PerishableProductDelivery-v0":
        return _perishable_instance(p)
    N, E = p["n_nodes"], p["n_edges"]
    weighted = p.get("weighted", True)
    n_graph = N - 1 if env_id == "DensestSubgraph-v0" else N       # densest_subgraph.py:59-65
    while True:
        edge_order = []
        adj = gnm_adjacency(n_graph, E, edge_order)
        if not _connected(adj, n_graph):
            continue
        if env_id == "TSP-v0":                                      # tsp.py:60-71
            if any(len(a) == 1 for a in adj):
                continue
            if not _connected(adj, n_graph, skip=0):
                continue
        break
    if env_id == "DensestSubgraph-v0":
        adj.append(dict())
    links = np.array([(u, v) for u in range(N) for v in adj[u]], dtype=np.int32).reshape(-1, 2)
    M = links.shape[0]
    ins = Instance(n_nodes=N, links=links, w64=np.ones(M, dtype=np.float64), edge_order=edge_order)
    rnd = np.random

    def matrix_delay(unweighted_lo, unweighted_hi, div):
        if weighted:
            delay = rnd.randint(3, 10, size=(N, N)) / 10.0
        else:
            delay = rnd.randint(unweighted_lo, unweighted_hi, size=(N, N)) / div
        lo = np.minimum(links[:, 0], links[:, 1])
        hi = np.maximum(links[:, 0], links[:, 1])
        return delay[lo, hi].astype(np.float64)  # d['delay'] = delay[u, v] with (u, v) in G.edges order => u < v

    if env_id in ("ShortestPath-v0", "LongestPath-v0"):
        ins.w64 = matrix_delay(10, 11, 10.0)                        # shortest_path.py:59-67
        s, t = rnd.choice(N, size=2, replace=False)                 # :74
        ins.src, ins.dest = int(s), int(t)
    elif env_id == "SteinerTree-v0":
        ins.w64 = matrix_delay(1, 2, 1.0)                           # steiner_tree.py:62-68
        d = rnd.choice(N, p["n_dests"] + 1, replace=False)          # :73
        ins.src, ins.dests = int(d[0]), d[1:].astype(np.int32)      # :89
    elif env_id == "TSP-v0":
        wmap = {}
        if p.get("spatial"):                                        # tsp.py:79-86
            xy = np.zeros((N, 2), dtype=np.float64)
            for v in range(N):
                xy[v, 0] = rnd.rand() * 10
                xy[v, 1] = rnd.rand() * 10
            for u, v in _undirected_edges(adj):
                wmap[(u, v)] = np.sqrt((xy[u, 0] - xy[v, 0]) ** 2 + (xy[u, 1] - xy[v, 1]) ** 2)
            ins.node_xy = xy
        else:                                                       # :88-93 one scalar draw per undirected edge
            for u, v in _undirected_edges(adj):
                wmap[(u, v)] = (rnd.randint(3, 10) / 10.0) if weighted else (rnd.randint(1, 2) / 1.0)
        ins.w64 = np.array([wmap[(u, v)] if (u, v) in wmap else wmap[(v, u)] for u, v in links], dtype=np.float64)
    elif env_id == "MaxIndependentSet-v0":                          # max_independent_set.py:53-57
        ins.node_cost = (rnd.randint(3, 10, size=N) / 10.0) if weighted else (rnd.randint(1, 2, size=N) / 1.0)
    elif env_id == "DensestSubgraph-v0":
        pass
    elif env_id == "MulticastRouting-v0":
        ins.w64 = matrix_delay(1, 2, 1.0)                           # multicast_routing.py:83-89
        ins.src = 0
        ins.dests = rnd.choice(np.arange(1, N), size=p["n_dests"], replace=False).astype(np.int32)  # :95
        ins.u01 = float(rnd.rand())                                 # :103 (max_distance itself needs the SSSP)
    elif env_id == "DistributionCenter-v0":
        ins.w64 = matrix_delay(1, 2, 1.0)                           # distribution_center.py:74-80
        ins.node_cost = rnd.randint(1, 4, size=N) / 1.0             # :82
        ins.dests = rnd.choice(N, size=p["target_count"], replace=False).astype(np.int32)  # :87
    else:
        raise KeyError(env_id)
    return ins
