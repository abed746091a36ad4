"""OPTIONAL host-side delegate for the three eval heuristics whose VALUE is defined by networkx's iteration
order: Kou's Steiner 2-approximation (steiner_tree.py:84-85), Christofides TSP
(tsp.py:114-117) and Ramsey clique-removal MIS (max_independent_set.py:62-65).  Their results depend on
Python set/dict iteration order inside networkx (tie-breaking among equal tenth-valued weights), so the only
way to return the reference's numbers is to run the same networkx routines on an nx.Graph rebuilt with the
same node / edge insertion order.  They run once per reset of an `is_eval_env` env, on the host, exactly
where the reference runs them; nothing on the step / mask / observation path touches this module.
Without networkx they return None and the env reports `heuristic_solution = nan` with a warning.
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
