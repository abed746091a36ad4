"""Static description of the 8 environments: ids, constructor kwargs/defaults and the per-env
`parenting` rules of the reference constructors, plus the layout table of
graph_envs/utils.py:32-73 (`get_env_info`)."""
import math
from dataclasses import dataclass


@dataclass(frozen=True)
class EnvSpec:
    kind: int
    node_f: int          # env-specific node columns (before the 5 structural ones)
    edge_f: int
    action_type: str     # "node" | "edge"
    step_w: str          # arithmetic of the edge weight inside step(): "f64" | "f32" | "none"
    uses_adj: bool       # adjacency bit-matrix resident on the device
    has_targets: bool
    has_node_cost: bool
    defaults: tuple      # ((kwarg, default), ...) after n_nodes, n_edges

    def heuristic_on_device(self, p):
        """Eval heuristics the GPU computes with the REFERENCE'S value: the tie-independent ones (Dijkstra
        for ShortestPath / LongestPath / SteinerTree n_dests=1, MST weight for n_dests=N-1) and Multicast's union of
        first-found shortest paths (networkx's pop order restated in csrc/ge_heuristics.cu)."""
        if self.kind in (0, 1, 6, 8):
            return True
        if self.kind == 2:
            return p["n_dests"] == 1 or p["n_dests"] == p["n_nodes"] - 1
        return False

    def heuristic_alternative(self, p):
        """Name of the labelled alternative heuristic the GPU computes where the reference's value is defined by networkx's
        set / dict iteration order (Kou, Christofides, Ramsey), else None.  Reported under
        info['heuristic_device'], never as info['heuristic_solution']."""
        if self.kind == 2 and not self.heuristic_on_device(p):
            return "steiner_shortest_path_heuristic"      # Takahashi-Matsuyama, 2-approximation like Kou
        if self.kind == 3:
            return "tsp_nearest_neighbour_walk"
        if self.kind == 4:
            return "mis_greedy_min_degree"
        return None


ENV_SPECS = {
    # graph_envs/shortest_path.py:23
    "ShortestPath-v0": EnvSpec(0, 2, 1, "node", "f64", True, False, False,
                               (("weighted", True), ("return_graph_obs", False), ("parenting", -1),
                                ("structural_features", True), ("is_eval_env", False))),
    # graph_envs/longest_path.py:26
    "LongestPath-v0": EnvSpec(1, 2, 1, "node", "f64", True, False, False,
                              (("weighted", True), ("return_graph_obs", False), ("is_eval_env", False), ("parenting", -1))),
    # graph_envs/steiner_tree.py:26
    "SteinerTree-v0": EnvSpec(2, 2, 2, "edge", "f32", False, True, False,
                              (("n_dests", 3), ("weighted", True), ("parenting", -1), ("is_eval_env", False))),
    # graph_envs/tsp.py:22
    "TSP-v0": EnvSpec(3, 4, 1, "node", "f64", True, False, False,
                      (("weighted", True), ("return_graph_obs", False), ("parenting", -1), ("spatial", False),
                       ("is_eval_env", False))),
    # graph_envs/max_independent_set.py:25
    "MaxIndependentSet-v0": EnvSpec(4, 2, 1, "node", "none", False, False, True,
                                    (("weighted", True), ("return_graph_obs", False), ("is_eval_env", False))),
    # graph_envs/densest_subgraph.py:25
    "DensestSubgraph-v0": EnvSpec(5, 1, 1, "node", "none", True, False, False,
                                  (("weighted", False), ("n_choices", -1), ("return_graph_obs", False),
                                   ("is_eval_env", False), ("parenting", -1))),
    # graph_envs/multicast_routing.py:31
    "MulticastRouting-v0": EnvSpec(6, 4, 2, "edge", "f32", False, True, False,
                                   (("n_dests", 3), ("weighted", True), ("max_distance", -1), ("parenting", 4),
                                    ("is_eval_env", False))),
    # graph_envs/distribution_center.py:29
    "DistributionCenter-v0": EnvSpec(7, 5, 1, "node", "f64", False, True, True,
                                     (("weighted", True), ("max_distance", 1), ("target_count", -1),
                                      ("return_graph_obs", False), ("is_eval_env", False), ("parenting", 2))),
    # graph_envs/perishable_product_delivery.py:26
    "PerishableProductDelivery-v0": EnvSpec(8, 16, 1, "node", "f64", True, False, False,
                                            (("n_products", 3), ("delivery_time", -1), ("weighted", True), ("return_graph_obs", False),
                                             ("is_eval_env", False), ("parenting", -1))),
}


def get_num_features():
    """graph_envs/feature_extraction.py:40-41"""
    return 5


def get_env_info(env_id):
    """graph_envs/utils.py:32-73: (node_f incl. structural, edge_f, action_type)."""
    if env_id not in ENV_SPECS:
        assert False, "Unknown env_id"
    s = ENV_SPECS[env_id]
    return s.node_f + get_num_features(), s.edge_f, s.action_type


def _density_edges(n_nodes):
    # longest_path.py:41-42, densest_subgraph.py:36-37, multicast_routing.py:52-53
    return int((n_nodes * (n_nodes - 1) // 2) * 0.30)


def check_ctor_args(env_id, n_nodes, n_edges, kwargs):
    """Applies the reference constructors' defaults and argument checks; returns the parameter dict."""
    spec = ENV_SPECS[env_id]
    p = dict(spec.defaults)
    for k, v in kwargs.items():
        if k not in p:
            raise TypeError("%s.__init__() got an unexpected keyword argument '%s'" % (env_id, k))
        p[k] = v
    p["n_nodes"] = int(n_nodes)
    par = p.get("parenting")
    if env_id == "ShortestPath-v0":
        assert par == -1, "Parenting is not available for shortest path"          # shortest_path.py:26
    elif env_id == "LongestPath-v0":
        assert par in [0, 1, 2, 3]                                                   # longest_path.py:29
        if n_edges == -1:
            n_edges = _density_edges(n_nodes)
    elif env_id == "SteinerTree-v0":
        assert par == -1, "Parenting not available for this environment"           # steiner_tree.py:29
    elif env_id == "TSP-v0":
        assert par in [1, 2], "Parenting must be either 1 or 2"                     # tsp.py:25
        if p["spatial"]:
            assert p["weighted"] == True, "Spatial TSP must be weighted"            # noqa: E712  tsp.py:26-27
    elif env_id == "DensestSubgraph-v0":
        assert par in [0, 1], "Parenting must be 0 or 1"                            # densest_subgraph.py:28
        assert p["weighted"] == False, "Weighted graphs not supported for this env"  # noqa: E712  :29
        if n_edges == -1:
            n_edges = _density_edges(n_nodes)
        if p["n_choices"] == -1:
            p["n_choices"] = n_nodes // math.e                                       # :38-39 (a float)
        p["n_choices"] = int(p["n_choices"]) if float(p["n_choices"]).is_integer() else -1
    elif env_id == "MulticastRouting-v0":
        if par not in [1, 2, 3, 4]:
            raise ValueError("Invalid parenting type")                              # multicast_routing.py:34-35
        if n_edges == -1:
            n_edges = _density_edges(n_nodes)
    elif env_id == "PerishableProductDelivery-v0":
        assert par in [1], "Parenting must be 1!"                                   # perishable_product_delivery.py:29
        assert p["n_products"] <= 5, "Max 5 products!"                              # :34
        if n_edges == -1:
            n_edges = _density_edges(n_nodes)                                        # :48-49
        # :51-58 -- the delivery-time range exists only for delivery_time == -1 (any other value makes the reference's
        # reset() fail on the missing attribute)
        assert p["delivery_time"] == -1, "PerishableProductDelivery: only delivery_time=-1 is usable in the reference"
        avg_degree = 2 * n_edges / n_nodes
        avg_dist = math.log(n_nodes) / math.log(avg_degree)
        if p["weighted"]:
            avg_dist = avg_dist * (0.3 + 1.0) / 2.0
        p["dt_mn"], p["dt_mx"] = avg_dist * 0.6, avg_dist * 1.4
    elif env_id == "DistributionCenter-v0":
        assert par in [1, 2]                                                         # distribution_center.py:32
        if p["target_count"] == -1:
            p["target_count"] = n_nodes // 5                                         # :42-43
    p["n_edges"] = int(n_edges)
    return p
