class Discrete:
    def __init__(self, n):
        self.n = n

class Box:
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.shape = low, high, shape

class GraphInstance:
    def __init__(self, nodes, edges, edge_links):
        self.nodes, self.edges, self.edge_links = nodes, edges, edge_links
