"""Layout contract of the flat observation (graph_envs/utils.py).

vectorize_graph   utils.py:87-88   [nodes.ravel | edges.ravel | edge_links.ravel] as float32
devectorize_graph utils.py:14-23   batched inverse: x[B,N,F], edge_features[B,2E,Fe], edge_index[B,2E,2]
get_env_info      utils.py:32-73   (node_f incl. the 5 structural columns, edge_f, action_type)

`to_pyg_graph` / `show_graph` are consumer-side (torch_geometric / matplotlib) and out of scope.
"""
import numpy as np

from .spec import get_env_info, get_num_features  # noqa: F401


class GraphInstance:
    """Attribute bag with the fields of gymnasium.spaces.GraphInstance (nodes, edges, edge_links)."""
    __slots__ = ("nodes", "edges", "edge_links")

    def __init__(self, nodes, edges, edge_links):
        self.nodes, self.edges, self.edge_links = nodes, edges, edge_links


def vectorize_graph(graph):
    return np.concatenate((graph.nodes.flatten(), graph.edges.flatten(), graph.edge_links.flatten()), dtype=np.float32)


def devectorize_graph(vector, env_id, **kwargs):
    """vector: [B, L] numpy array or torch tensor (host or device)."""
    bs = vector.shape[0]
    node_f, edge_f, _ = get_env_info(env_id)
    n, m = kwargs["n_nodes"], 2 * kwargs["n_edges"]
    p1 = n * node_f
    p2 = p1 + m * edge_f
    x = vector[:, :p1].reshape(bs, n, node_f)
    edge_features = vector[:, p1:p2].reshape(bs, m, edge_f)
    edge_index = vector[:, p2:].reshape(bs, m, 2)
    edge_index = edge_index.long() if hasattr(edge_index, "long") else edge_index.astype(np.int64)
    return x, edge_features, edge_index


def graph_from_obs(obs, env_id, n_nodes, n_edges):
    """Single flat obs -> GraphInstance (what info['graph_obs'] holds in the reference)."""
    x, ef, ei = devectorize_graph(np.asarray(obs)[None, :], env_id, n_nodes=n_nodes, n_edges=n_edges)
    return GraphInstance(x[0], ef[0], ei[0])
