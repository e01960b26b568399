"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (the parity ORACLE) of the reference hot path:
k-mer de Bruijn graph construction + non-branching-path (NBP) compaction + NBP origins of
fpusan/SuperPang.  Citations are `file:line` under the reference's `src/superpang/`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module.
The product path (`superpang_b200/`) never does; it fails loudly if its CUDA library is missing.

Parity status: PINNED.  The reference ships no golden vectors, so this restatement is
pinned against the reference itself: `oracle/refrun.py` executes the unmodified reference
`Assembler.__init__` + its own `is_NBP_init` / `reconstruct_sequence` / `get_ori2cov_onevertex`
in the build container, `tests/test_oracle_vs_reference.py` compares the two field by field, and
`oracle/make_golden.py` froze reference outputs into `tests/golden/*.json`, which this module must
reproduce wherever it runs (see `tests/test_oracle_golden.py`).  One item is pinned by reasoning
only: the rotation chosen for pure-cycle components depends on graph-tool's neighbour order
(`lib/Assembler.py:311-316`), reproduced here as "out-edges in insertion order, then in-edges".

The restatement is deliberately structured like the reference (per-sequence hashing, one serial
dict pass, undirected simple graph, per-component serial chain walk) but uses numpy for the
byte/integer bulk work so that multi-million-k-mer inputs finish in seconds to minutes.
"""
import re
from collections import defaultdict

import numpy as np

# ------------------------------------------------------------------------------------------------
# input: lib/utils.py:5-30, :52-59
# ------------------------------------------------------------------------------------------------
_AMBIG = {ord(a): ord("N") for a in "RYKMSWBDHV"}          # utils.py:52
_ALLOWED = set("ACTGN")                                     # utils.py:53
_RE_REMOVE = re.compile(r"[\s.-]+")                         # utils.py:10


def read_fasta_split(path, gap_size=1):
    """`read_fasta(fasta, ambigs='as_Ns', Ns='split', gap_size=1)` (utils.py:5-30; called Assembler.py:94).

    Upper-case, U->T, IUPAC ambiguity -> N, illegal characters raise, records are split at every
    gap of `gap_size` Ns, pieces stripped of Ns, named `{name}_Nsplit_{i}` only when the record
    contained a gap; duplicated names raise.  Returns an insertion-ordered dict.
    """
    gap = "N" * gap_size
    out = {}
    txt = open(path).read().strip().lstrip(">")
    for rec in txt.split("\n>"):
        name, seq = rec.split("\n", 1)
        name = name.split(" ")[0]
        seq = _RE_REMOVE.sub("", seq).upper().replace("U", "T")
        seq = seq.translate(_AMBIG)
        bad = set(seq) - _ALLOWED
        if bad:
            raise ValueError(f'Illegal chars in seq "{name}": {bad}')
        if name in out:
            raise Exception(f'Sequence "{name}" is duplicated in your input file')
        if gap not in seq:
            out[name] = seq
        else:
            i = 0
            for piece in seq.split(gap):
                piece = piece.strip("N")
                if piece:
                    out[f"{name}_Nsplit_{i}"] = piece
                    i += 1
    return out


_COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def reverse_complement(seq):
    """cutils.pyx:16-28 (A<->T, C<->G, N->N, reversed)."""
    if isinstance(seq, str):
        return seq.encode().translate(_COMP)[::-1].decode()
    return seq.translate(_COMP)[::-1]


# ------------------------------------------------------------------------------------------------
# hashing: lib/Compressor.pyx:32-58, lib/Assembler.py:978-992
# ------------------------------------------------------------------------------------------------
_CODE = {65: 0, 67: 1, 71: 2, 84: 3}


def compress(kmer):
    """`Compressor.compress` (Compressor.pyx:32-58): k//4 bytes of 2-bit codes (A=0,C=1,G=2,T=3, first
    base in the top bits of each byte) followed by the k%4 trailing bases as raw ASCII."""
    if isinstance(kmer, str):
        kmer = kmer.encode()
    k = len(kmer)
    com, rem = k // 4, k % 4
    out = bytearray(com + rem)
    for i in range(com):
        t = kmer[4 * i:4 * i + 4]
        out[i] = (_CODE[t[0]] << 6) | (_CODE[t[1]] << 4) | (_CODE[t[2]] << 2) | _CODE[t[3]]
    for i in range(rem):
        out[com + i] = kmer[4 * com + i]
    return bytes(out)


def canonical_keys(seq, k):
    """`seq2hashes` (Assembler.py:978-992): key(i) = max(compress(fwd k-mer), compress(rc k-mer)).

    compress() is injective and order-preserving w.r.t. plain lexicographic order of the base string
    (A<C<G<T both as 2-bit codes and as ASCII), so max over compressed bytes == compress(max over raw
    k-mers); the dedup dictionary therefore partitions instances identically when keyed by the raw
    maximum.  (`tests/test_oracle_golden.py::test_compress_order_equivalence` checks this against the
    reference's compiled Compressor.)  Yields the raw canonical k-mer bytes and a flag fwd-is-canonical.
    """
    s = seq.encode() if isinstance(seq, str) else seq
    rc = reverse_complement(s)
    n = len(s)
    for i in range(n - k + 1):
        h = s[i:i + k]
        ri = n - i - k
        rh = rc[ri:ri + k]
        yield (h, True) if h >= rh else (rh, False)


def compress_vertices(sp):
    """vtools.pyx:18-49: sort, then run-length encode into closed intervals.  Reproduces the
    `last = cs = 0` initial-state quirk: if the smallest id is 1 the first interval starts at 0."""
    vs = np.sort(np.asarray(sp, dtype=np.int64))
    res = []
    first = True
    last = 0
    cs = 0
    # vectorised equivalent of the reference loop
    brk = np.nonzero(np.diff(vs) > 1)[0]
    starts = np.concatenate([[vs[0]], vs[brk + 1]])
    ends = np.concatenate([vs[brk], [vs[-1]]])
    if starts[0] - 0 <= 1:       # `v - last > 1` is false for the first element when v in {0, 1}
        starts[0] = 0
    del res, first, last, cs
    return np.stack([starts, ends], axis=1).astype(np.uint32)


# ------------------------------------------------------------------------------------------------
# DBG build: lib/Assembler.py:80-256
# ------------------------------------------------------------------------------------------------
def build_dbg(seqDict, k):
    """Vertices = distinct canonical k-mers numbered by first appearance (Assembler.py:186-195);
    `vertex2coords[v] = (ref, idx)` of that first appearance, `ref` counting skipped short sequences
    too (:150-153); one edge per consecutive instance pair (:199-201), rows sorted, np.unique (:247-249);
    `seqLimits` = first/last vertex of every sequence (:206-211); `seqPaths[name]` =
    compress_vertices(sp) (:220)."""
    hash2vertex = {}
    coords = []
    sps = {}
    seqLimits = set()
    edge_chunks = []
    n_inst = 0
    for ref, (name, seq) in enumerate(seqDict.items()):
        if len(seq) <= k:                                   # :151
            continue
        nk = len(seq) - k + 1
        n_inst += nk
        sp = np.empty(nk, dtype=np.uint32)
        nv = len(hash2vertex)
        for idx, (h, _) in enumerate(canonical_keys(seq, k)):
            v = hash2vertex.get(h)
            if v is None:
                v = nv
                hash2vertex[h] = v
                coords.append((ref, idx))
                nv += 1
            sp[idx] = v
        seqLimits.add(int(sp[0]))
        seqLimits.add(int(sp[-1]))
        if nk > 1:
            edge_chunks.append(np.stack([sp[:-1], sp[1:]], axis=1))
        sps[name] = sp
    vertex2coords = np.asarray(coords, dtype=np.uint32).reshape(-1, 2)
    if edge_chunks:
        edges = np.concatenate(edge_chunks)
        edges.sort(axis=1)
        edges = np.unique(edges, axis=0)
    else:
        edges = np.zeros((0, 2), dtype=np.uint32)
    seqPaths = {name: compress_vertices(sp) for name, sp in sps.items()}
    return dict(vertex2coords=vertex2coords, edges=edges.astype(np.uint32), seqPaths=seqPaths,
                seqLimits=seqLimits, sps=sps, n_inst=n_inst)


class _Adj:
    """Undirected simple-graph adjacency in graph-tool order (out-edges in insertion order, then
    in-edges); an undirected self-loop appears twice (Assembler.py:251-252)."""

    def __init__(self, edges, nv):
        E = edges.shape[0]
        s = edges[:, 0].astype(np.int64)
        t = edges[:, 1].astype(np.int64)
        owner = np.concatenate([s, t])
        other = np.concatenate([t, s])
        part = np.concatenate([np.zeros(E, np.int64), np.ones(E, np.int64)])
        eid = np.concatenate([np.arange(E), np.arange(E)])
        order = np.lexsort((eid, part, owner))
        self.nbr = other[order]
        self.ptr = np.concatenate([[0], np.cumsum(np.bincount(owner, minlength=nv))])
        self.nv = nv

    def neighbors(self, v):
        return self.nbr[self.ptr[v]:self.ptr[v + 1]]


def _components(edges, nv):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    m = coo_matrix((np.ones(edges.shape[0], np.int8), (edges[:, 0], edges[:, 1])), shape=(nv, nv))
    _, lab = connected_components(m, directed=False)
    _, first = np.unique(lab, return_index=True)           # label = rank of the component's lowest vertex id
    order = np.argsort(first)
    remap = np.empty_like(order)
    remap[order] = np.arange(order.size)
    return remap[lab]


# ------------------------------------------------------------------------------------------------
# NBP collection: lib/Assembler.py:269-363, :995-1118
# ------------------------------------------------------------------------------------------------
class _Kmers:
    def __init__(self, seqDict, vertex2coords, k):
        self.seqs = [s for s in seqDict.values()]
        self.v2c = vertex2coords
        self.k = k

    def kmer(self, v):                                      # vertex2kmer, :1062-1069
        ref, idx = self.v2c[v]
        return self.seqs[ref][idx:idx + self.k]

    def extend(self, k1, v2, force=False):                  # extendKmer, :1092-1107
        k2x = self.kmer(v2)
        hits = [k2 for k2 in (k2x, reverse_complement(k2x)) if k2[:-1] == k1[1:]]
        if force:
            assert len(hits) == 1, "degenerate input (Q8): extension is not unique"
        else:
            assert len(hits) <= 1, "degenerate input (Q8): extension is not unique"
        return hits[0] if hits else None


def is_nbp_init(v, adj, seqLimits, K):
    """is_NBP_init, :995-1035."""
    succs = set(int(x) for x in adj.neighbors(v)) - {v}     # :1013
    if len(succs) != 2:
        return True
    if v in seqLimits:                                      # :1017 fake-extender rule
        kmers = set()
        k1x = K.kmer(v)
        for s in succs:
            for k1 in (k1x, reverse_complement(k1x)):
                if K.extend(k1, s):
                    kmers.add(k1)
        if len(kmers) == 1:
            return True
    return False


def reconstruct_sequence(path, K):
    """reconstruct_sequence + edge2kmer + seq_from_kmers, :1038-1059, :1072-1089, :1110-1118."""
    k1x = K.kmer(path[0])
    start = []
    for k1 in (k1x, reverse_complement(k1x)):
        k2 = K.extend(k1, path[1])
        if k2:
            start.extend([k1, k2])
    assert len(start) == 2, "degenerate input (Q8): first edge orientation is not unique"   # :1088
    out = [start[0], start[1][-1]]
    cur = start[1]
    for v in path[2:]:
        cur = K.extend(cur, v, force=True)
        out.append(cur[-1])
    return "".join(out)


def collect_nbps(seqDict, dbg, k):
    """Per connected component (size desc, ties by label; :271-273): init detection (:302), cycle
    breaking (:304-325), serial chain walk (:327-339), cycle closing (:341-343), terminal-duplicate
    filter (:354), sequence reconstruction (:357)."""
    v2c, edges = dbg["vertex2coords"], dbg["edges"]
    nv = v2c.shape[0]
    adj = _Adj(edges, nv)
    K = _Kmers(seqDict, v2c, k)
    comp = _components(edges, nv) if nv else np.zeros(0, np.int64)
    labels, sizes = np.unique(comp, return_counts=True)
    order = [c for c, s in sorted(zip(labels, sizes), key=lambda x: x[1], reverse=True)]
    seqLimits = dbg["seqLimits"]
    # distinct non-self degree for every vertex (vectorised pre-filter for the init test)
    e = edges[edges[:, 0] != edges[:, 1]]
    deg = np.bincount(np.concatenate([e[:, 0], e[:, 1]]).astype(np.int64), minlength=nv)
    out = []
    for ic, c in enumerate(order):
        vs = np.nonzero(comp == c)[0]
        inits = set(int(v) for v in vs[deg[vs] != 2])
        for v in vs[deg[vs] == 2]:
            v = int(v)
            if v in seqLimits and is_nbp_init(v, adj, seqLimits, K):
                inits.add(v)
        isCycle = False
        cut = None
        if not inits:
            if min(len(set(adj.neighbors(v).tolist())) for v in vs) == 2:        # :309
                for v in vs:
                    nb = adj.neighbors(v)
                    if len(set(nb.tolist())) == 2:                               # :312
                        inits.add(int(v))
                        inits.add(int(nb[0]))
                        isCycle = True
                        cut = {int(v), int(nb[0])}
                        break

        def nbrs(x):
            r = [int(y) for y in adj.neighbors(x)]
            if cut is not None and x in cut:
                other = (cut - {x}).pop() if len(cut) == 2 else x
                r = [y for y in r if y != other]
            return r

        NBPs = []
        for ini in inits:                                                        # :327
            for n in set(nbrs(ini)):
                NBP = [ini, n]
                last = ini
                while NBP[-1] not in inits:
                    cur = NBP[-1]
                    s = [x for x in nbrs(cur) if x != last and x != cur][0]      # :334
                    last = cur
                    NBP.append(s)
                NBPs.append(tuple(NBP))
        if isCycle:                                                              # :341-343
            assert len(NBPs) == 2
            NBPs = [p + (p[0],) for p in NBPs]
        NBPs = [p for p in NBPs if p[0] != p[1] and p[-1] != p[-2]]              # :354
        for p in NBPs:
            out.append((p, reconstruct_sequence(p, K), ic, isCycle))
    return out, comp


# ------------------------------------------------------------------------------------------------
# origins: lib/Assembler.py:1212-1262, :296, :944-946
# ------------------------------------------------------------------------------------------------
def nbp_origins(nbps, seqDict, dbg, comp):
    """`get_ori2cov_onevertex` over `seqPathsThisComp` with an empty bubble map, then goodOris.

    vertex2origins(v) = names whose interval list contains v (:1257-1262, vtools.pyx:89-103) -- the
    interval lists include the compress_vertices leading-zero quirk.  Only sequences whose
    `sp[0][0]` lies in the NBP's component are considered (:296).  NBP with >=3 vertices: union over
    the interior vertices (:1226-1240); 2 vertices: names containing both (cov 1.0 survives :944; the
    max-cov name is added :945).  Names cut at `_Nsplit_` (:946)."""
    names = list(dbg["seqPaths"])
    nv = dbg["vertex2coords"].shape[0]
    vs_l, ss_l = [], []
    seq_comp = {}
    for si, n in enumerate(names):
        iv = dbg["seqPaths"][n].astype(np.int64)
        seq_comp[si] = int(comp[iv[0, 0]])
        lens = iv[:, 1] - iv[:, 0] + 1
        starts = np.repeat(iv[:, 0] - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
        vs_l.append(starts + np.arange(lens.sum()))
        ss_l.append(np.full(int(lens.sum()), si, dtype=np.int64))
    if vs_l:
        vs = np.concatenate(vs_l)
        ss = np.concatenate(ss_l)
        o = np.lexsort((ss, vs))
        vs, ss = vs[o], ss[o]
    else:
        vs = ss = np.zeros(0, np.int64)
    ptr = np.searchsorted(vs, np.arange(nv + 1))
    short = [n.split("_Nsplit_")[0] for n in names]
    res = []
    for p, seq, ic, cyc in nbps:
        vcomp = int(comp[p[0]])
        if len(p) > 2:
            inner = np.asarray(p[1:-1], dtype=np.int64)
            lo, hi = ptr[inner], ptr[inner + 1]
            cnt = hi - lo
            idx = np.repeat(lo - np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt) + np.arange(cnt.sum())
            cand = np.unique(ss[idx])
            good = {int(s) for s in cand if seq_comp[int(s)] == vcomp}
        else:
            a = set(ss[ptr[p[0]]:ptr[p[0] + 1]].tolist())
            b = set(ss[ptr[p[1]]:ptr[p[1] + 1]].tolist())
            cov = defaultdict(float)
            for s in a:
                if seq_comp[s] == vcomp:
                    cov[s] += 0.5
            for s in b:
                if seq_comp[s] == vcomp:
                    cov[s] += 0.5
            good = {s for s, c in cov.items() if c >= 1}
            # :945 "add at least the origin with the highest cov" -- first max in dict (= seqPaths) order
            if cov:
                best = sorted(sorted(cov), key=lambda s: cov[s], reverse=True)[0]
                good.add(best)
        res.append((p, seq, tuple(sorted({short[s] for s in good})), ic, cyc))
    return res


def run_oracle(fasta_or_dict, k=301, with_origins=True):
    seqDict = read_fasta_split(fasta_or_dict) if isinstance(fasta_or_dict, str) else dict(fasta_or_dict)
    dbg = build_dbg(seqDict, k)
    nbps, comp = collect_nbps(seqDict, dbg, k)
    if with_origins:
        nbps = nbp_origins(nbps, seqDict, dbg, comp)
    else:
        nbps = [(p, s, (), ic, cyc) for p, s, ic, cyc in nbps]
    return dict(
        k=k,
        names=list(seqDict),
        n_inst=dbg["n_inst"],
        vertex2coords=dbg["vertex2coords"],
        edges=dbg["edges"],
        seqPaths=dbg["seqPaths"],
        seqLimits=sorted(dbg["seqLimits"]),
        ncomp=int(comp.max()) + 1 if comp.size else 0,
        nbps=nbps,
        _sps=dbg["sps"],
        _comp=comp,
    )


def nbp_orientation_counts(nbps, seqs):
    """for every raw NBP (vertices, sequence, ...) the number of input sequences that contain its
    string, as the reference accumulates it into nv2rightOrientation (lib/Assembler.py:448-453) from
    get_seqPaths_NBPs (:1121-1139).  The pre-filter there, `vertex_overlap(compress_vertices(NBP), seqPath) == refLen`
    (:1135), is implied by the substring test (:1137) except for the compress_vertices leading-zero quirk
    (lib/vtools.pyx:29-45): a vertex list whose smallest vertex is 1 is recorded as holding vertex 0 as well, and so
    is the path of every sequence that holds vertex 1, which makes the overlap refLen + 1 -- such an NBP is never
    counted."""
    out = []
    for nbp in nbps:
        verts, seq = nbp[0], nbp[1]
        if min(verts) == 1:
            out.append(0)
            continue
        out.append(sum(1 for s in seqs if seq in s))
    return out


def bubble_candidates(nbps):
    """the reference's `path_limits` (lib/Assembler.py:728-734): NBPs grouped by (first vertex, last
    vertex); only groups with more than one path are bubble candidates."""
    from collections import defaultdict
    groups = defaultdict(list)
    for nbp in nbps:
        verts = tuple(nbp[0])
        groups[(verts[0], verts[-1])].append(verts)
    return [g for g in groups.values() if len(g) > 1]


def seq_paths_nbps(nbps, seqdict):
    """{sequence name: indices of the raw NBPs whose string the sequence contains}, i.e. the value of
    the reference's get_seqPaths_NBPs (lib/Assembler.py:1121-1139) for every sequence, with the same leading-zero quirk
    as in nbp_orientation_counts; sequences without any NBP are left out (:449)."""
    out = {}
    for name, s in seqdict.items():
        mine = [i for i, nbp in enumerate(nbps) if min(nbp[0]) != 1 and nbp[1] in s]
        if mine:
            out[name] = mine
    return out
