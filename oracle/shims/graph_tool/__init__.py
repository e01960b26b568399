"""
TEST INFRASTRUCTURE ONLY -- minimal stand-in for the `graph_tool` package.

The reference (fpusan/SuperPang, `src/superpang/lib/Assembler.py:14-15`) imports graph-tool,
which is not installed in this image.  This stub provides the subset that `Assembler.__init__`
(`Assembler.py:251-255`) and `Assembler.run` (`Assembler.py:271-334`, `:376-384`, `:418-424`, `:468-472`, `:500`,
`:537-568`, `:635`, `:682-700`, `:767-779`, `:969-975`) touch, so that the *unmodified* reference code can
be executed as the parity oracle (`oracle/refrun.py`) and for the drop-in integration test
(`oracle/dropin.py`).  Checklist: SURVEY.md appendix A.

Semantics implemented:
  * `Graph(directed=False).add_edge_list(uint32[E,2])`: vertex count = max id + 1; `get_out_neighbors(v)` lists the
    neighbours through edges in which `v` is the source (edge-insertion order) followed by those in which `v`
    is the target; an undirected self-loop is listed twice; honours the edge filter.
  * `Graph()` (directed): `add_vertex`, `add_edge` (parallel edges allowed), `get_out/in/all_neighbors`,
    `vertices()` / `edges()` / `num_vertices()` / `num_edges()`, all honouring the vertex filter, which is a LIVE
    view of the property map handed to `set_vertex_filter` (as in graph-tool).
  * property maps index by vertex / edge descriptor or int, expose `get_array()`, and ITERATE over the visible
    vertices only (`zip(g.vertices(), label_components(g)[0])`, `Assembler.py:540`).
  * `label_components(g, directed=False)`: components of the visible sub-graph, numbered in order of their
    lowest vertex id.

Nothing under `superpang_b200/` imports this module.
"""
import numpy as np
from . import topology  # noqa: F401


class _PropertyMap:
    def __init__(self, arr, graph=None):
        self.a = arr
        self._g = graph            # set for vertex maps: iteration follows the graph's vertex filter

    def get_array(self):
        return self.a

    def __getitem__(self, k):
        return self.a[int(k)]

    def __setitem__(self, k, v):
        self.a[int(k)] = v

    def __iter__(self):
        if self._g is not None and self._g._vfilt is not None:
            return iter(self.a[self._g._visible_mask()])
        return iter(self.a)

    def __len__(self):
        return len(self.a)


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
        self._edges = np.zeros((0, 2), dtype=np.int64)      # static edge list (add_edge_list)
        self._adj_ptr = None
        self._efilt = None
        self._vfilt = None                                  # live _PropertyMap (or None)
        self._vfilt_inv = False
        self._dyn_out, self._dyn_in, self._dyn_edges = [], [], []   # dynamic graphs (add_vertex / add_edge)
        self._dyn_out_e, self._dyn_in_e = [], []                    # edge index of every adjacency entry (edge filter)
        self.vp = _PropDict()
        self.ep = _PropDict()

    # -- construction -------------------------------------------------------------------
    def add_edge_list(self, edges):
        if self._directed:                                   # condense-edges.py:64: directed graph, vertices added before
            for a, b in edges:
                while self._nv <= max(int(a), int(b)):
                    self.add_vertex()
                self.add_edge(a, b)
            return
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

    def add_vertex(self):
        self._dyn_out.append([])
        self._dyn_in.append([])
        self._dyn_out_e.append([])
        self._dyn_in_e.append([])
        self._nv += 1
        return self._nv - 1

    def add_edge(self, a, b):
        a, b = int(a), int(b)
        self._dyn_out[a].append(b)
        self._dyn_in[b].append(a)
        self._dyn_out_e[a].append(len(self._dyn_edges))
        self._dyn_in_e[b].append(len(self._dyn_edges))
        self._dyn_edges.append((a, b))
        return _Edge(a, b, len(self._dyn_edges) - 1)

    # -- filters -------------------------------------------------------------------------
    def _visible_mask(self):
        m = np.asarray(self._vfilt.a, dtype=bool)
        return ~m if self._vfilt_inv else m

    def _vis(self, v):
        if self._vfilt is None:
            return True
        x = bool(self._vfilt.a[v])
        return (not x) if self._vfilt_inv else x

    def _evis(self, e):
        return self._efilt is None or e >= len(self._efilt) or bool(self._efilt[e])

    def set_vertex_filter(self, prop, inverted=False):
        self._vfilt = prop
        self._vfilt_inv = inverted

    def set_edge_filter(self, prop, inverted=False):
        if prop is None:
            self._efilt = None
            return
        f = np.asarray(prop.get_array(), dtype=bool)
        self._efilt = ~f if inverted else f.copy()

    def clear_filters(self):
        self._efilt = None
        self._vfilt = None
        self._vfilt_inv = False

    # -- queries ------------------------------------------------------------------------
    def num_vertices(self):
        if self._vfilt is not None:
            return int(np.count_nonzero(self._visible_mask()))
        return self._nv

    def num_edges(self):
        if self._adj_ptr is not None:
            if self._efilt is not None:
                return int(np.count_nonzero(self._efilt))
            return self._edges.shape[0]
        return sum(1 for i, (a, b) in enumerate(self._dyn_edges) if self._vis(a) and self._vis(b) and self._evis(i))

    def vertices(self):
        for v in range(self._nv):
            if self._vis(v):
                yield v

    def get_out_neighbors(self, v):
        v = int(v)
        if self._adj_ptr is not None:
            lo, hi = self._adj_ptr[v], self._adj_ptr[v + 1]
            nbr = self._adj_nbr[lo:hi]
            if self._efilt is not None:
                nbr = nbr[self._efilt[self._adj_eid[lo:hi]]]
            return nbr
        return np.asarray([x for x, e in zip(self._dyn_out[v], self._dyn_out_e[v]) if self._vis(x) and self._evis(e)], dtype=np.int64)

    def get_in_neighbors(self, v):
        v = int(v)
        if self._adj_ptr is not None:
            return self.get_out_neighbors(v)
        return np.asarray([x for x, e in zip(self._dyn_in[v], self._dyn_in_e[v]) if self._vis(x) and self._evis(e)], dtype=np.int64)

    def get_all_neighbors(self, v):
        v = int(v)
        if self._adj_ptr is not None:
            return self.get_out_neighbors(v)
        return np.asarray([x for x, e in zip(self._dyn_out[v] + self._dyn_in[v], self._dyn_out_e[v] + self._dyn_in_e[v])
                           if self._vis(x) and self._evis(e)], dtype=np.int64)

    def edge(self, a, b):
        a, b = int(a), int(b)
        lo, hi = self._adj_ptr[a], self._adj_ptr[a + 1]
        for nb, e in zip(self._adj_nbr[lo:hi], self._adj_eid[lo:hi]):
            if nb == b:
                return _Edge(int(self._edges[e, 0]), int(self._edges[e, 1]), int(e))
        return None

    def edges(self):
        if self._adj_ptr is not None:
            for e in range(self._edges.shape[0]):
                if self._efilt is None or self._efilt[e]:
                    yield _Edge(int(self._edges[e, 0]), int(self._edges[e, 1]), e)
        else:
            for i, (a, b) in enumerate(self._dyn_edges):
                if self._vis(a) and self._vis(b) and self._evis(i):
                    yield _Edge(a, b, i)

    def _visible_edge_array(self):
        if self._adj_ptr is not None:
            e = self._edges
            return e[self._efilt] if self._efilt is not None else e
        if not self._dyn_edges:
            return np.zeros((0, 2), dtype=np.int64)
        e = np.asarray(self._dyn_edges, dtype=np.int64)
        if self._efilt is not None:
            e = e[np.asarray(self._efilt[:e.shape[0]], dtype=bool)]
        if self._vfilt is not None:
            m = self._visible_mask()
            e = e[m[e[:, 0]] & m[e[:, 1]]]
        return e

    # -- properties -----------------------------------------------------------------------
    def new_vertex_property(self, typ):
        return _PropertyMap(np.zeros(self._nv, dtype=_TYPES[typ]), graph=self)

    def new_edge_property(self, typ):
        n = self._edges.shape[0] if self._adj_ptr is not None else len(self._dyn_edges)
        return _PropertyMap(np.zeros(n, dtype=_TYPES[typ]))
