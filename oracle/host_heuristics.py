"""TEST INFRASTRUCTURE (oracle side; nothing under graphenvs_b200/ imports this).  Host restatements of the reference's
eval heuristics, used to pin the recorded reference values in tests/golden/heuristics.json:

* Kou's Steiner 2-approximation (steiner_tree.py:84-85), Christofides TSP (tsp.py:114-117) and Ramsey clique-removal MIS
  (max_independent_set.py:62-65): their VALUES are defined by Python set/dict iteration order inside networkx, so the
  only way to reproduce them is to run the same networkx routines on an nx.Graph rebuilt with the same node / edge
  insertion order (needs networkx: build container only).  The product ships labelled alternatives computed on the
  device instead (csrc/ge_heuristics.cu, info['heuristic_device']).
* MulticastRouting's union of first-found shortest paths (multicast_routing.py:107-115): a literal restatement of nx
  `_dijkstra_multisource`; the device kernel (csrc/ge_heuristics.cu:heur_multicast_kernel) is tested against it.
"""


def _nx():
    try:
        import networkx as nx
        return nx
    except Exception:
        return None


def available():
    return _nx() is not None


def _graph(nx, n, edge_order, attr=None, wmap=None):
    G = nx.Graph()
    G.add_nodes_from(range(n))            # gnm_random_graph / complete_graph: nodes first, then edges in draw order
    G.add_edges_from(edge_order)
    if attr is not None:
        for u, v, d in G.edges(data=True):
            d[attr] = wmap[(u, v)] if (u, v) in wmap else wmap[(v, u)]
    return G


def steiner_kou(n, edge_order, wmap, terminals):
    """sum of 'delay' over nx steiner_tree(G, terminals, weight='delay', method='kou').edges(); terminals = the
    reference's np.random.choice array (source first)."""
    nx = _nx()
    if nx is None:
        return None
    G = _graph(nx, n, edge_order, "delay", wmap)
    T = nx.algorithms.approximation.steinertree.steiner_tree(G, terminals, weight="delay", method="kou")
    return float(sum([G[u][v]["delay"] for u, v in T.edges()]))


def tsp_christofides(n, edge_order, wmap):
    nx = _nx()
    if nx is None:
        return None
    G = _graph(nx, n, edge_order, "weight", wmap)
    cycle = nx.approximation.traveling_salesman_problem(G, weight="weight", cycle=True)
    total = 0
    for i in range(len(cycle) - 1):
        total += G[cycle[i]][cycle[i + 1]]["weight"]
    return float(total)


def mis_ramsey(n, edge_order):
    nx = _nx()
    if nx is None:
        return None
    G = _graph(nx, n, edge_order)
    return float(len(nx.approximation.maximum_independent_set(G)))


def multicast_union_of_paths(n, links, w64, src, dests):
    """MulticastRouting eval heuristic (multicast_routing.py:107-115): total weight of the union of the
    FIRST-FOUND shortest paths src -> each destination.  The value depends on networkx's tie order, so the
    search restates nx `_dijkstra_multisource` literally (nx:algorithms/shortest_paths/weighted.py:853-881):
    heap of (dist, insertion counter, node), neighbours in adjacency insertion order (= row order of `links`), a
    predecessor is replaced only by a strictly shorter path; paths are rebuilt from the first predecessor; the edge
    set and its sum use the same Python set / list operations as the reference."""
    from heapq import heappop, heappush
    from itertools import count, islice
    adj = [[] for _ in range(n)]
    wm = {}
    for (u, v), w in zip(links.tolist(), w64):
        adj[u].append(v)
        wm[(u, v)] = w          # numpy float64 like the reference's G[u][v]['delay']: Python >= 3.12 sums *float* objects with
                                # compensated summation, numpy scalars with plain left-to-right adds
    dist, seen, pred = {}, {src: 0}, {}
    c = count()
    fringe = [(0, next(c), src)]
    while fringe:
        d_v, _, v = heappop(fringe)
        if v in dist:
            continue
        dist[v] = d_v
        for u in adj[v]:
            vu = d_v + wm[(v, u)]
            if u in dist:
                continue
            if u not in seen or vu < seen[u]:
                seen[u] = vu
                heappush(fringe, (vu, next(c), u))
                pred[u] = v
    paths = {src: [src]}
    for v in islice(dist, 1, None):
        paths[v] = paths[pred[v]] + [v]
    edges = set()
    for d in dests:
        path = paths[int(d)]
        for u, v in zip(path[:-1], path[1:]):
            edges.add((u, v))
    return sum([wm[(u, v)] for u, v in edges])


def reference_heuristic(env_id, p, ins):
    """The reference's info['heuristic_solution'] for a host-generated instance where its value is tie-dependent
    (None when this module has nothing to add or networkx is missing for the three networkx-defined ones)."""
    N = ins.n_nodes
    wmap = {(int(u), int(v)): float(w) for (u, v), w in zip(ins.links.tolist(), ins.w64)}
    if env_id == "MulticastRouting-v0":
        return float(multicast_union_of_paths(N, ins.links, ins.w64, 0, ins.dests))
    if env_id == "SteinerTree-v0" and 1 < p["n_dests"] < N - 1:
        import numpy as np
        return steiner_kou(N, ins.edge_order, wmap, np.concatenate([[ins.src], ins.dests]))
    if env_id == "TSP-v0":
        return tsp_christofides(N, ins.edge_order, wmap)
    if env_id == "MaxIndependentSet-v0" and not p.get("weighted", True):
        return mis_ramsey(N, ins.edge_order)
    return None
