"""TEST INFRASTRUCTURE ONLY -- `graph_tool.topology.label_components` stand-in (see package docstring)."""
import numpy as np


def label_components(g, directed=False):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from . import _PropertyMap

    n = g.num_vertices()
    e = g._edges
    if g._efilt is not None:
        e = e[g._efilt]
    m = coo_matrix((np.ones(e.shape[0], dtype=np.int8), (e[:, 0], e[:, 1])), shape=(n, n))
    _, lab = connected_components(m, directed=False)
    # renumber in order of lowest vertex id (graph-tool labels components in discovery order)
    _, first = np.unique(lab, return_index=True)
    order = np.argsort(first)
    remap = np.empty_like(order)
    remap[order] = np.arange(order.size)
    lab = remap[lab].astype(np.int32)
    hist = np.bincount(lab)
    return _PropertyMap(lab), hist
