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


def flat_obs_sections(env_id, n_nodes, n_edges):
    """(start, stop, shape) of the three sections of the flat observation: node matrix, edge matrix, edge list."""
    node_f, edge_f, _ = get_env_info(env_id)
    m = 2 * n_edges
    a = n_nodes * node_f
    b = a + m * edge_f
    return (0, a, (n_nodes, node_f)), (a, b, (m, edge_f)), (b, b + 2 * m, (m, 2))


def vectorize_graph(graph):
    """GraphInstance -> float32 vector [nodes | edges | edge_links], each row-major (utils.py:87-88)."""
    parts = (np.asarray(graph.nodes), np.asarray(graph.edges), np.asarray(graph.edge_links))
    out = np.empty(sum(p.size for p in parts), dtype=np.float32)
    pos = 0
    for p in parts:
        out[pos:pos + p.size] = p.reshape(-1)
        pos += p.size
    return out


def devectorize_graph(vector, env_id, **kwargs):
    """Batched inverse of vectorize_graph (utils.py:14-23).  vector: [B, L] numpy array or torch tensor (host or
    device); returns (x [B,N,F], edge_features [B,2E,Fe], edge_index [B,2E,2] as int64)."""
    batch = vector.shape[0]
    out = []
    for lo, hi, shape in flat_obs_sections(env_id, kwargs["n_nodes"], kwargs["n_edges"]):
        out.append(vector[:, lo:hi].reshape((batch,) + shape))
    x, edge_features, edge_index = out
    edge_index = edge_index.long() if hasattr(edge_index, "long") else edge_index.astype(np.int64)
    return x, edge_features, edge_index


def graph_from_obs(obs, env_id, n_nodes, n_edges):
    """Single flat obs -> GraphInstance (what info['graph_obs'] holds in the reference)."""
    x, ef, ei = devectorize_graph(np.asarray(obs)[None, :], env_id, n_nodes=n_nodes, n_edges=n_edges)
    return GraphInstance(x[0], ef[0], ei[0])
