"""
TEST INFRASTRUCTURE ONLY -- minimal stand-in for the `graph_tool` package.

The reference (fpusan/SuperPang, `src/superpang/lib/Assembler.py:14-15`) imports graph-tool,
which is not installed in this image.  This stub provides exactly the subset that the
hot path (`Assembler.__init__`, `Assembler.py:251-255`, and the NBP-collection head of
`Assembler.run`, `Assembler.py:271-334`) touches, so that the *unmodified* reference code can
be executed as the parity oracle (see `oracle/refrun.py`).

Semantics implemented:
  * `Graph(directed=False).add_edge_list(uint32[E,2])`: vertex count = max id + 1.
  * `get_out_neighbors(v)`: for an undirected graph, the neighbours through edges in which `v`
    is the source (edge-insertion order) followed by those in which `v` is the target; an
    undirected self-loop is listed twice (graph-tool behaviour); honours the edge filter.
  * `label_components(g, directed=False)`: components numbered in order of their lowest vertex id.

Nothing under `superpang_b200/` imports this module.
"""
import numpy as np
from . import topology  # noqa: F401


class _PropertyMap:
    def __init__(self, arr):
        self.a = arr

    def get_array(self):
        return self.a

    def __getitem__(self, k):
        return self.a[int(k)]

    def __setitem__(self, k, v):
        self.a[int(k)] = v

    def __iter__(self):
        return iter(self.a)


class _PropDict:
    """`g.vp.name = prop` / `g.vp.name` attribute container."""
    pass


class _Edge:
    __slots__ = ("s", "t", "idx")

    def __init__(self, s, t, idx):
        self.s, self.t, self.idx = s, t, idx

    def source(self):
        return self.s

    def target(self):
        return self.t

    def __int__(self):
        return self.idx

    def __hash__(self):
        return hash(self.idx)

    def __eq__(self, other):
        return isinstance(other, _Edge) and other.idx == self.idx


_TYPES = {"bool": np.bool_, "int16_t": np.int16, "int32_t": np.int32, "int": np.int32}


class Graph:
    def __init__(self, directed=True):
        self._directed = directed
        self._nv = 0
        self._edges = np.zeros((0, 2), dtype=np.int64)
        self._adj_ptr = None
        self._efilt = None
        self.vp = _PropDict()
        self.ep = _PropDict()

    # -- construction -------------------------------------------------------------------
    def add_edge_list(self, edges):
        edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        self._edges = edges
        self._nv = int(edges.max()) + 1 if edges.size else 0
        self._build_adj()

    def _build_adj(self):
        E = self._edges.shape[0]
        s, t = self._edges[:, 0], self._edges[:, 1]
        eid = np.arange(E, dtype=np.int64)
        # out part (v is source) then in part (v is target); each in edge-insertion order
        owner = np.concatenate([s, t])
        other = np.concatenate([t, s])
        part = np.concatenate([np.zeros(E, np.int64), np.ones(E, np.int64)])
        eids = np.concatenate([eid, eid])
        order = np.lexsort((eids, part, owner))
        self._adj_nbr = other[order]
        self._adj_eid = eids[order]
        counts = np.bincount(owner, minlength=self._nv)
        self._adj_ptr = np.concatenate([[0], np.cumsum(counts)])

    # -- queries ------------------------------------------------------------------------
    def num_vertices(self):
        return self._nv

    def num_edges(self):
        if self._efilt is not None:
            return int(np.count_nonzero(self._efilt))
        return self._edges.shape[0]

    def get_out_neighbors(self, v):
        v = int(v)
        lo, hi = self._adj_ptr[v], self._adj_ptr[v + 1]
        nbr = self._adj_nbr[lo:hi]
        if self._efilt is not None:
            nbr = nbr[self._efilt[self._adj_eid[lo:hi]]]
        return nbr

    get_all_neighbors = get_out_neighbors

    def edge(self, a, b):
        a, b = int(a), int(b)
        lo, hi = self._adj_ptr[a], self._adj_ptr[a + 1]
        for nb, e in zip(self._adj_nbr[lo:hi], self._adj_eid[lo:hi]):
            if nb == b:
                return _Edge(int(self._edges[e, 0]), int(self._edges[e, 1]), int(e))
        return None

    def edges(self):
        for e in range(self._edges.shape[0]):
            if self._efilt is None or self._efilt[e]:
                yield _Edge(int(self._edges[e, 0]), int(self._edges[e, 1]), e)

    # -- properties / filters -------------------------------------------------------------
    def new_vertex_property(self, typ):
        return _PropertyMap(np.zeros(self._nv, dtype=_TYPES[typ]))

    def new_edge_property(self, typ):
        return _PropertyMap(np.zeros(self._edges.shape[0], dtype=_TYPES[typ]))

    def set_edge_filter(self, prop, inverted=False):
        if prop is None:
            self._efilt = None
            return
        f = np.asarray(prop.get_array(), dtype=bool)
        self._efilt = ~f if inverted else f.copy()

    def clear_filters(self):
        self._efilt = None
