"""Helpers for the -m gpu tests: build BatchedGraphEnv batches from golden cases / seeded instances."""
import hashlib

import numpy as np

import golden_util as gu


LANE_KINDS = ("ShortestPath-v0", "LongestPath-v0", "TSP-v0", "MaxIndependentSet-v0", "DensestSubgraph-v0")


def has_fast_path(env_id, n_nodes, parenting=None):
    """True when the default dispatch is NOT the full-recompute warp-per-env kernel (lane-per-env for
    N <= 64 node-action kinds, incremental-mask kernels for SteinerTree / Multicast p>=2 / MIS), i.e.
    when force_warp=True exercises a different code path worth testing."""
    if n_nodes <= 64 and env_id in LANE_KINDS:
        return True
    if env_id in ("ShortestPath-v0", "DensestSubgraph-v0", "LongestPath-v0", "TSP-v0"):
        return True   # group-per-env kernels (64 < N <= 1024)
    if env_id == "PerishableProductDelivery-v0":
        return n_nodes <= 64   # lane-per-env kernel vs the warp-per-env one
    if env_id in ("SteinerTree-v0", "MaxIndependentSet-v0", "DistributionCenter-v0"):
        return True   # (DistributionCenter: exact distance automaton vs the fp64 search)
    return env_id == "MulticastRouting-v0" and (parenting is None or parenting >= 2)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]


def group_key(m):
    kw = m["kwargs"]
    return (m["env_id"], m["N"], m["M"], kw.get("parenting", None), kw.get("n_dests", None), kw.get("n_choices", None),
            kw.get("target_count", None), kw.get("max_distance", None), kw.get("weighted", True), kw.get("spatial", False))


def grouped_cases():
    groups = {}
    for m, r in gu.all_cases():
        groups.setdefault(group_key(m), []).append((m, r))
    return groups


def instance_from_case(m, r):
    from graphenvs_b200.instances import Instance
    k = gu.instance_kwargs(m, r)
    return Instance(n_nodes=k["N"], links=k["links"], w64=k["w64"], src=k.get("src", 0), dest=k.get("dest", 0),
                    dests=k.get("dests"), node_cost=k.get("node_cost"), node_xy=k.get("node_xy"),
                    max_distance=k.get("max_distance"), heuristic=k.get("heuristic", 0.0), features=k["features"])


def batch_from_cases(cases, **extra):
    from graphenvs_b200 import BatchedGraphEnv
    m0 = cases[0][0]
    kw = dict(m0["kwargs"])
    n_nodes, n_edges = kw.pop("n_nodes"), kw.pop("n_edges")
    env = BatchedGraphEnv(m0["env_id"], len(cases), n_nodes, n_edges, structural_features=True, **kw, **extra)
    env.load_instances([instance_from_case(m, r) for m, r in cases])
    return env


def oracle_from_instance(env_id, ins, params, features=None):
    from oracle import oracle as orc
    kind = orc.KINDS[env_id]
    md = params.get("max_distance", 0.0) if env_id == "DistributionCenter-v0" else (ins.max_distance or 0.0)
    return orc.OracleEnv(kind, ins.n_nodes, ins.links, ins.w64, parenting=params.get("parenting", -1),
                         features=features, src=ins.src, dest=ins.dest,
                         n_dests=params.get("n_dests", 0), n_choices=params.get("n_choices", 0) or 0,
                         max_distance=md, heuristic=ins.heuristic or 0.0, dests=ins.dests,
                         node_cost=ins.node_cost, node_xy=ins.node_xy)
