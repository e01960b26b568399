"""TEST INFRASTRUCTURE ONLY -- `graph_tool.topology.label_components` stand-in (see package docstring)."""
import numpy as np


def label_components(g, directed=False):
    """Components of the VISIBLE sub-graph (vertex / edge filters honoured), numbered in order of their lowest
    vertex id.  Hidden vertices get label -1 in the underlying array; iterating the returned map yields the
    labels of the visible vertices only."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from . import _PropertyMap

    n = g._nv
    e = g._visible_edge_array()
    m = coo_matrix((np.ones(e.shape[0], dtype=np.int8), (e[:, 0], e[:, 1])), shape=(n, n))
    _, lab = connected_components(m, directed=False)
    vis = g._visible_mask() if g._vfilt is not None else np.ones(n, dtype=bool)
    out = np.full(n, -1, dtype=np.int32)
    if vis.any():
        lv = lab[vis]
        _, first, inv = np.unique(lv, return_index=True, return_inverse=True)
        order = np.argsort(first)                       # labels in order of first (lowest-id) visible vertex
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        out[vis] = rank[inv].astype(np.int32)
    hist = np.bincount(out[out >= 0]) if (out >= 0).any() else np.zeros(0, dtype=np.int64)
    return _PropertyMap(out, graph=g), hist


def all_circuits(g):
    """Elementary circuits of the visible directed graph as vertex lists (condense-edges.py:111 only sorts them by length
    and reads their first vertex)."""
    import networkx as nx
    G = nx.DiGraph()
    G.add_nodes_from(g.vertices())
    G.add_edges_from((e.source(), e.target()) for e in g.edges())
    return [list(c) for c in nx.simple_cycles(G)]
