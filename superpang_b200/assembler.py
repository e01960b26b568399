"""
Host-side mirror of the reference's `Assembler` for the accelerated path.

`Assembler(fasta, ksize, threads, diskdb=None, debug=False, gap_size=1)` has the reference's
constructor signature (`src/superpang/lib/Assembler.py:80`) and fills the attributes its host tail
reads -- `ksize`, `seqDict`, `ref2name`, `vertex2coords`, `seqPaths`, `seqLimits`,
`seqLimitsPerSeq`, `includedSeqs` (`lib/Assembler.py:85-96`, `:194`, `:206-220`, `:246`) -- from the
CUDA library instead of the Python dict loop.  `collect_nbps()` replaces the NBP-collection head of
`Assembler.run` (`lib/Assembler.py:269-363`): per connected component, in the reference's processing
order, the raw NBP vertex tuples, `NBP2seq` sorted by (len, seq) descending (`:362`), the cycle flag
(`:305`) and the per-NBP origin names after the goodOris rule (`:944-946`).

Everything numeric happens on the GPU behind the C-ABI (`include/superpang_b200.h`); this module only
converts device results into the Python containers the reference's downstream code expects.
"""
import re
from collections import namedtuple

import numpy as np

from ._capi import Engine, SPB_NONE

_AMBIG = {ord(a): ord("N") for a in "RYKMSWBDHV"}
_ALLOWED = frozenset("ACTGN")
_RE_REMOVE = re.compile(r"[\s.-]+")

Component = namedtuple("Component", ["label", "n_vertices", "NBPs", "NBP2seq", "isCycle", "origins"])


def read_fasta(fasta, gap_size=1):
    """Same behaviour as the reference's `read_fasta(fasta, ambigs='as_Ns', Ns='split', gap_size)`
    (`lib/utils.py:5-30`, `:52-59`): strip whitespace/`.`/`-`, upper-case, U->T, IUPAC ambiguity -> N,
    illegal characters raise ValueError, split at every run of `gap_size` Ns, strip Ns from the pieces,
    name them `{name}_Nsplit_{i}` only if the record had a gap; duplicated names raise."""
    gap = "N" * gap_size
    seqDict = {}
    with open(fasta) as fh:
        txt = fh.read().strip().lstrip(">")
    for rec in txt.split("\n>"):
        name, seq = rec.split("\n", 1)
        name = name.split(" ")[0]
        seq = _RE_REMOVE.sub("", seq).upper().replace("U", "T").translate(_AMBIG)
        bad = set(seq) - _ALLOWED
        if bad:
            raise ValueError(f'Illegal chars in seq "{name}": {bad}')
        if name in seqDict:
            raise Exception(f'Sequence "{name}" is duplicated in your input file')
        if gap not in seq:
            seqDict[name] = seq
        else:
            i = 0
            for piece in seq.split(gap):
                piece = piece.strip("N")
                if piece:
                    seqDict[f"{name}_Nsplit_{i}"] = piece
                    i += 1
    return seqDict


def concat_sequences(seqs):
    """list of str / bytes / uint8 arrays -> (uint8[T], uint64[n+1])."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.empty(int(off[-1]), dtype=np.uint8)
    for s, a in zip(seqs, off[:-1]):
        a = int(a)
        if isinstance(s, np.ndarray):
            buf[a:a + len(s)] = s
        else:
            b = s.encode() if isinstance(s, str) else s
            buf[a:a + len(b)] = np.frombuffer(b, dtype=np.uint8)
    return buf, off


class Assembler:
    engine_factory = None          # tests inject the kernel-logic emulator here; the product default is Engine(device) = CUDA

    def __init__(self, fasta, ksize, threads=1, diskdb=None, debug=False, gap_size=1, *, engine=None, device=0, dist_group=None,
                 distributed=False):
        """Build the de Bruijn graph on the GPU.

        `threads` and `diskdb` (`--lowmem`) are accepted for drop-in compatibility and ignored: the
        fork pool and the RocksDB k-mer store (`lib/Assembler.py:109-117`, `:148`) are replaced by HBM.
        `gap_size != 1` (raw k-mer strings with Ns as keys, `lib/Assembler.py:173-180`) is not covered
        by the 2-bit GPU path and is refused."""
        if gap_size != 1:
            raise NotImplementedError("the B200 path only covers --gap-size 1 (2-bit k-mers)")
        self.ksize = int(ksize)
        self.debug = debug
        if isinstance(fasta, dict):
            self.seqDict = {n: (s if isinstance(s, str) else bytes(s).decode()) for n, s in fasta.items()}
        else:
            self.seqDict = read_fasta(fasta, gap_size=gap_size)
        if any("N" in s for s in self.seqDict.values()):
            raise ValueError("sequences still contain N after splitting")
        self.ref2name = list(self.seqDict)
        self.includedSeqs = sum(1 for s in self.seqDict.values() if len(s) > self.ksize)
        if engine is None:
            engine = type(self).engine_factory() if type(self).engine_factory is not None else Engine(device=device)
        self._eng = engine
        buf, off = concat_sequences(list(self.seqDict.values()))
        self._seq_off = off
        self._eng.load(buf, off, self.ksize)
        # multi-GPU: one process per GPU (torch.distributed initialised by the caller), k-mers hash-partitioned over the ranks
        self.counts = self._eng.build_dbg_distributed(dist_group) if (distributed or dist_group is not None) else self._eng.build_dbg()
        self._compacted = False
        self._cache = {}

    def _full(self):
        """Multi-GPU: graph build and compaction ran sliced over the ranks (Engine._mg_sliced_tail).  The reference's host
        tail is one process and reads whole arrays, so the first accessor that needs them has this rank recompute the
        sliced passes over the whole input (Engine.complete: no communication; counts, origins, components and the NBP
        table are complete on every rank anyway)."""
        if getattr(self._eng, "mg_sliced", False):
            self.counts = self._eng.complete()
            self._compacted = True

    # ---- DBG-build state consumed by the reference's host tail ------------------------------
    @property
    def vertex2coords(self):
        self._full()
        if "v2c" not in self._cache:
            self._cache["v2c"] = self._eng.vertex2coords()
        return self._cache["v2c"]

    @property
    def edges(self):
        self._full()
        if "edges" not in self._cache:
            self._cache["edges"] = self._eng.edges()
        return self._cache["edges"]

    def _limits(self):
        if "lim" not in self._cache:
            self._cache["lim"] = self._eng.seqlimits()
        return self._cache["lim"]

    @property
    def seqLimits(self):
        lim = self._limits()
        return set(int(x) for x in lim.ravel() if x != SPB_NONE)

    @property
    def seqLimitsPerSeq(self):
        lim = self._limits()
        return {n: {int(a), int(b)} for n, (a, b) in zip(self.ref2name, lim) if a != SPB_NONE}

    @property
    def seqPaths(self):
        self._full()
        if "sp" not in self._cache:
            off, iv = self._eng.seq_intervals()
            self._cache["sp"] = {n: iv[int(off[i]):int(off[i + 1])] for i, n in enumerate(self.ref2name)
                                 if off[i + 1] > off[i]}
        return self._cache["sp"]

    def instance_vertices(self, name):
        """The per-sequence `sp` array of `lib/Assembler.py:183-197` (vertex id of every k-mer)."""
        self._full()
        if "inst" not in self._cache:
            self._cache["inst"] = self._eng.instance_vertices()
        i = self.ref2name.index(name)
        a, b = int(self._seq_off[i]), int(self._seq_off[i + 1])
        return self._cache["inst"][a:b - self.ksize + 1] if b - a > self.ksize else np.zeros(0, np.uint32)

    # ---- NBP collection ---------------------------------------------------------------------
    def compact(self):
        self._full()
        if not self._compacted:
            self.counts = self._eng.compact()
            self._compacted = True
        return self.counts

    def raw_nbps(self):
        """[(vertex tuple, sequence str, origin-name tuple, component label, isCycle)] for every raw NBP."""
        self.compact()
        voff, verts, comp, cyc = self._eng.nbps()
        soff, seqs = self._eng.nbp_seqs()
        ooff, oidx = self._eng.nbp_origins()
        short = [n.split("_Nsplit_")[0] for n in self.ref2name]           # lib/Assembler.py:946
        # one bulk conversion each instead of one numpy call per NBP (the containers themselves -- a tuple of vertex ids and
        # a str per NBP -- are what the reference's host tail works on and cannot be avoided here)
        vl, sbytes = verts.tolist(), seqs.tobytes()
        vo, so, oo, ol = voff.tolist(), soff.tolist(), ooff.tolist(), oidx.tolist()
        cl, cy = comp.tolist(), cyc.tolist()
        names_of = {}                                                     # origin-index lists repeat a lot: name tuples are cached
        out = []
        for i in range(len(cl)):
            key = tuple(ol[oo[i]:oo[i + 1]])
            o = names_of.get(key)
            if o is None:
                o = names_of[key] = tuple(sorted({short[j] for j in key}))
            out.append((tuple(vl[vo[i]:vo[i + 1]]), sbytes[so[i]:so[i + 1]].decode(), o, cl[i], bool(cy[i])))
        return out

    def nbp_orientation_counts(self):
        """{vertex tuple of a raw NBP: number of input sequences that contain its string} -- what the reference
        accumulates into nv2rightOrientation (`lib/Assembler.py:448-453`) from `get_seqPaths_NBPs` (`:1121-1139`)."""
        self.compact()
        voff, verts, _, _ = self._eng.nbps()
        counts = self._eng.nbp_orientation_counts()
        return {tuple(verts[int(voff[i]):int(voff[i + 1])].tolist()): int(counts[i]) for i in range(len(counts))}

    def seq_paths_nbps(self):
        """{sequence name: set of raw NBPs (vertex tuples) the sequence contains in its own orientation} -- the value of
        `get_seqPaths_NBPs` for every sequence (`lib/Assembler.py:1121-1139`); names without any NBP are left out, as
        the reference drops them (`:449`)."""
        self.compact()
        voff, verts, _, _ = self._eng.nbps()
        off, ids = self._eng.seq_nbps()
        tup = {}
        out = {}
        for s, name in enumerate(self.ref2name):
            mine = ids[int(off[s]):int(off[s + 1])]
            if len(mine):
                for i in mine:
                    if int(i) not in tup:
                        tup[int(i)] = tuple(verts[int(voff[i]):int(voff[i + 1])].tolist())
                out[name] = {tup[int(i)] for i in mine}
        return out

    def nbp_successors(self):
        """{vertex tuple of a raw NBP: tuple of the vertex tuples of its successors} -- NBP2 follows NBP1 when the last k bases
        of NBP1 are the first k bases of NBP2, the edge rule of the reference's NBP graph (`lib/Assembler.py:365-384`);
        NBPs without successors are left out.  From `spb_get_nbp_graph` (a k-base string is a (vertex, orientation) pair there)."""
        self.compact()
        voff, verts, _, _ = self._eng.nbps()
        off, succ, _rev = self._eng.nbp_graph()
        tup = [tuple(verts[int(voff[i]):int(voff[i + 1])].tolist()) for i in range(len(voff) - 1)]
        return {tup[i]: tuple(tup[int(j)] for j in succ[int(off[i]):int(off[i + 1])]) for i in range(len(tup)) if off[i + 1] > off[i]}

    def reconstruct_sequences(self, paths):
        """`reconstruct_sequence` (`lib/Assembler.py:1038-1118`) for a list of vertex paths (e.g. joined NBPs)."""
        return self._eng.path_sequences(paths, self.ksize)

    def path_origins(self, paths, bubble_vertex2origins=None, nbp_ids=False):
        """Origin names (sorted, `_Nsplit_` suffixes cut as at `lib/Assembler.py:946`) of paths that are concatenations of
        whole NBPs -- `get_ori2cov_onevertex` + goodOris for the joined paths of the host tail.  paths: vertex tuples, or raw
        NBP ids (nbp_ids=True); bubble_vertex2origins: {vertex: origin names} as the host tail keeps it (`:922`)."""
        self.compact()
        short = [n.split("_Nsplit_")[0] for n in self.ref2name]
        bmap = None
        if bubble_vertex2origins:
            index = {n: i for i, n in enumerate(self.ref2name)}
            bmap = {int(v): [index[n] for n in names] for v, names in bubble_vertex2origins.items() if names}
        got = self._eng.path_origins(paths, nbp_ids=nbp_ids, bubble_vertex2origins=bmap)
        return [tuple(sorted({short[j] for j in ids})) for ids in got]

    def bubble_candidates(self):
        """Groups (lists of vertex tuples) of raw NBPs that share first and last vertex -- `path_limits` entries with more
        than one path, `lib/Assembler.py:728-734`."""
        self.compact()
        voff, verts, _, _ = self._eng.nbps()
        g = self._eng.bubble_groups()
        groups = {}
        for i in np.nonzero(g != 0xFFFFFFFF)[0]:
            groups.setdefault(int(g[i]), []).append(tuple(verts[int(voff[i]):int(voff[i + 1])].tolist()))
        return list(groups.values())

    def collect_nbps(self):
        """Per component, in the reference's order (size descending, ties by label; `lib/Assembler.py:271-273`)."""
        raw = self.raw_nbps()
        labels = self._eng.components()
        sizes = np.bincount(labels, minlength=int(self.counts["n_comp"]))
        order = [c for c, s in sorted(zip(range(len(sizes)), sizes), key=lambda x: x[1], reverse=True)]
        by = {c: [] for c in order}
        for r in raw:
            by[r[3]].append(r)
        comps = []
        for c in order:
            rows = sorted(by[c], key=lambda r: (len(r[1]), r[1]), reverse=True)        # lib/Assembler.py:362
            NBP2seq = {r[0]: r[1] for r in rows}
            comps.append(Component(label=c, n_vertices=int(sizes[c]), NBPs=list(NBP2seq), NBP2seq=NBP2seq,
                                   isCycle=any(r[4] for r in rows), origins={r[0]: set(r[2]) for r in rows}))
        return comps

    def run(self, minlen, mincov, bubble_identity_threshold, genome_assignment_threshold, max_threads, debug=False):
        """The reference's `Assembler.run` (`lib/Assembler.py:260`): dict[(scaffold, i)] -> Contig.  DBG, NBPs, NBP sequences and
        the per-sequence NBP sets come from the device; the host tail (`lib/Assembler.py:365-963`) is the installed
        reference's own code (see `superpang_b200/dropin.py`)."""
        from . import dropin
        return dropin.run(self, minlen, mincov, bubble_identity_threshold, genome_assignment_threshold, max_threads, debug)

    # ---- canonical result for parity tests ----------------------------------------------------
    def to_result(self):
        raw = self.raw_nbps()
        return dict(
            k=self.ksize,
            names=list(self.seqDict),
            n_inst=int(self.counts["n_inst"]),
            vertex2coords=self.vertex2coords,
            edges=self.edges,
            seqPaths=self.seqPaths,
            seqLimits=sorted(self.seqLimits),
            ncomp=int(self.counts["n_comp"]),
            nbps=raw,
        )
