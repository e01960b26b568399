#!/usr/bin/env python3
"""
TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference hot path as the parity oracle.

Only usable where `/root/reference` exists (the build container).  It imports the reference's
`superpang.lib.Assembler` (with the Cython helpers compiled by `oracle/build_ref.py` and the
`graph_tool` / `mappy` / `speedict` import shims under `oracle/shims/`) and drives:

  * `Assembler.__init__`                       (reference `lib/Assembler.py:80-256`)   -- DBG build
  * the NBP-collection head of `Assembler.run` (reference `lib/Assembler.py:269-363`)  -- restated
    below 1:1 as a driver that calls the reference's OWN classmethods `is_NBP_init`,
    `reconstruct_sequence` (`:995-1059`) and `get_ori2cov_onevertex` (`:1212-1248`) through the
    reference's own `multimap` (`:59-74`); the goodOris rule is `:944-946`.

It is used (a) to validate the independent restatement in `oracle/dbg_oracle.py` and (b) to
generate the golden vectors committed under `tests/golden/` (`oracle/make_golden.py`).
Nothing in the product path, the `-m gpu` tests, `smoke()` or `bench.py` imports this file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("SPB_REFERENCE", "/root/reference")


def available():
    """The reference can be run: either its tree is present (build container) or `build_ref.stage()` left a runnable
    copy under oracle/_ref/superpang (which travels to the GPU box)."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import build_ref
    return os.path.isdir(os.path.join(REFERENCE, "src", "superpang", "lib")) or build_ref.staged()


def reference_root():
    """Directory to put on sys.path so that `import superpang` finds the unmodified reference."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import build_ref
    if os.path.isdir(os.path.join(REFERENCE, "src", "superpang", "lib")):
        build_ref.stage(REFERENCE)
    if not build_ref.staged():
        raise RuntimeError("reference not present and not staged (run oracle/build_ref.py in the build container)")
    return os.path.dirname(build_ref.STAGED)


def import_reference():
    """Import the reference Assembler class (from the staged copy, oracle/_ref/superpang)."""
    root = reference_root()
    for p in (os.path.join(HERE, "shims"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    from superpang.lib.Assembler import Assembler
    return Assembler


def run_reference(fasta, ksize=301, threads=1, debug=True, with_origins=True, quiet=True, extras=True, timing=None):
    """Run the reference DBG build + NBP collection on `fasta`; return the canonical result dict.
    timing (dict, optional) receives the wall seconds of the three parts SURVEY 8(d) names: `init_s` = Assembler.__init__
    (`lib/Assembler.py:80-256`), `nbp_s` = the NBP-collection head of run() (`:269-363`), `origins_s` = the origin lookup
    (`:1212-1262`).  extras=False skips the 8(f) golden material (orientation counts, joined paths)."""
    Assembler = import_reference()
    import graph_tool as gt
    import io
    import contextlib
    import time as _time

    T = timing if timing is not None else {}
    T.update(init_s=0.0, nbp_s=0.0, origins_s=0.0)
    sink = io.StringIO() if quiet else sys.stdout
    with contextlib.redirect_stdout(sink):
        _t0 = _time.perf_counter()
        A = Assembler(fasta, ksize, threads, None, debug, 1)
        T["init_s"] = _time.perf_counter() - _t0
        _t0 = _time.perf_counter()
        # ---- head of Assembler.run, reference lib/Assembler.py:269-363 (driver restatement) ----
        A.DBG.vp.comp = gt.topology.label_components(A.DBG, directed=False)[0]            # :271
        comps, sizes = np.unique(A.DBG.vp.comp.get_array(), return_counts=True)              # :272
        comps = [c for c, s in sorted(zip(comps, sizes), key=lambda x: x[1], reverse=True)]  # :273
        nbps_out = []
        for ic, c in enumerate(comps):
            vs = np.where(A.DBG.vp.comp.get_array() == np.int16(c))[0]                       # :285
            vset = set(vs.tolist())
            seqPathsThisComp = {n: sp for n, sp in A.seqPaths.items() if int(sp[0][0]) in vset}  # :296
            NBPs = []
            badEdges = set()
            flags = A.multimap(A.is_NBP_init, threads, vs, A.DBG, A.seqLimits, A.vertex2coords,
                               A.ref2name, A.seqDict, A.ksize)                               # :302
            inits = {int(v) for v, f in zip(vs, flags) if f}
            isCycle = False
            if not inits:                                                                    # :307
                if min(len(set(A.DBG.get_out_neighbors(v))) for v in vs) == 2:               # :309
                    for v in vs:
                        neighbors = A.DBG.get_out_neighbors(v)
                        if len(set(neighbors)) == 2:                                         # :312
                            inits.add(int(v))
                            inits.add(int(neighbors[0]))
                            isCycle = True
                            badEdges.add(A.DBG.edge(v, neighbors[0]))
                            badEdges.add(A.DBG.edge(neighbors[0], v))
                            break
            if badEdges:                                                                     # :322
                for e in A.DBG.edges():
                    A.DBG.ep.efilt[e] = e not in badEdges
                A.DBG.set_edge_filter(A.DBG.ep.efilt)
            for ini in inits:                                                                # :327
                for n in set(sorted(A.DBG.get_out_neighbors(ini))):
                    NBP = [ini, int(n)]
                    last = ini
                    while True:
                        if NBP[-1] in inits:
                            break
                        s = [x for x in A.DBG.get_out_neighbors(NBP[-1]) if x != last and x != NBP[-1]][0]
                        last = NBP[-1]
                        NBP.append(int(s))
                    NBPs.append(tuple(NBP))
            if isCycle:                                                                      # :341
                assert len(NBPs) == 2
                NBPs = [NBP + (NBP[0],) for NBP in NBPs]
            for NBP in NBPs:
                assert len(NBP) > 1                                                          # :347
            NBPs = [NBP for NBP in NBPs if NBP[0] != NBP[1] and NBP[-1] != NBP[-2]]          # :354
            seqs = A.multimap(A.reconstruct_sequence, threads, NBPs, A.vertex2coords, A.ref2name,
                              A.seqDict, A.ksize)                                            # :357
            T["nbp_s"] += _time.perf_counter() - _t0
            _t0 = _time.perf_counter()
            if with_origins:
                from collections import defaultdict
                o2c = A.multimap(A.get_ori2cov_onevertex, threads, NBPs, seqPathsThisComp,
                                 defaultdict(list), A.seqLimitsPerSeq, debug, chunksize=10)  # :922
            else:
                o2c = [None] * len(NBPs)
            T["origins_s"] += _time.perf_counter() - _t0
            _t0 = _time.perf_counter()
            if badEdges:
                A.DBG.clear_filters()
            for NBP, seq, ori2cov in zip(NBPs, seqs, o2c):
                if ori2cov is None:
                    oris = ()
                else:                                                                        # :944-946
                    good = {o for o, cov in ori2cov.items() if cov >= 1}
                    good.add(list(sorted(ori2cov, key=lambda o: ori2cov[o], reverse=True))[0])
                    oris = tuple(sorted({o.split('_Nsplit_')[0] for o in good}))
                nbps_out.append((tuple(int(x) for x in NBP), seq, oris, ic, isCycle))

        # ---- orientation counts: the reference's own get_seqPaths_NBPs (:1121-1139) over all NBP pairs, counted as at
        #      :448-453 (nv2rightOrientation); -1 where the reference defines no pair (vertex set not shared by exactly two NBPs)
        from collections import defaultdict as _dd
        vertex2NBP = {i: n[0] for i, n in enumerate(nbps_out)}
        NBP2seq = {n[0]: n[1] for n in nbps_out}
        groups = _dd(list)
        for i, n in enumerate(nbps_out):
            groups[frozenset(n[0])].append(i)
        psets = {fs: tuple(nvs) for fs, nvs in groups.items() if extras and len(nvs) == 2 and len(NBP2seq) == len(nbps_out)}
        orient = [-1] * len(nbps_out)
        orient_sets = {}
        if psets:
            names_ = list(A.seqPaths)
            sets_ = A.multimap(A.get_seqPaths_NBPs, 1, names_, A.seqPaths, psets, NBP2seq, vertex2NBP, A.seqDict)
            for nvs in psets.values():
                for nv in nvs:
                    orient[nv] = 0
            for spGS in sets_:
                for nv in spGS:
                    orient[nv] += 1
            orient_sets = {name: sorted(int(nv) for nv in spGS) for name, spGS in zip(names_, sets_) if spGS}

        # ---- joined paths: NBP1 + NBP2[1:] for NBP-graph edges (:365-384), sequences by the reference's own
        #      reconstruct_sequence (:1038-1118) -- golden vectors for spb_path_sequences
        joined = []
        by_start = _dd(list)
        for n in nbps_out:
            by_start[n[1][:ksize]].append(n)
        for n1 in (nbps_out if extras else ()):
            for n2 in by_start[n1[1][-ksize:]]:
                if len(joined) < 40 and n1[0][-1] == n2[0][0] and len(set(n1[0]) & set(n2[0])) == 1:
                    joined.append(tuple(n1[0]) + tuple(n2[0][1:]))
        joined_seqs = A.multimap(A.reconstruct_sequence, 1, joined, A.vertex2coords, A.ref2name, A.seqDict, A.ksize) if joined else []
        joined_oris = []
        if joined and with_origins:                                                          # :922, :941-946
            for o2c_ in A.multimap(A.get_ori2cov_onevertex, 1, joined, A.seqPaths, _dd(list), A.seqLimitsPerSeq, debug, chunksize=10):
                good = {o for o, cov in o2c_.items() if cov >= 1}
                good.add(list(sorted(o2c_, key=lambda o: o2c_[o], reverse=True))[0])
                joined_oris.append(tuple(sorted({o.split('_Nsplit_')[0] for o in good})))
        else:
            joined_oris = [()] * len(joined)
        # ---- the same origin rule with a NON-EMPTY bubble_vertex2origins (:922, :1236): golden vectors for
        #      spb_path_origins_ex.  The map is made up deterministically from the paths themselves (a popped bubble leaves
        #      exactly such vertex -> origin-name lists behind); every raw NBP and joined path is then asked about.
        bubble_map, bubble_oris = {}, []
        if extras and with_origins:
            all_paths = [tuple(n[0]) for n in nbps_out] + list(joined)
            names_all = list(A.seqDict)
            for i, p_ in enumerate(all_paths):
                if i % 3 == 0 and names_all:
                    v_ = int(p_[len(p_) // 2] if len(p_) > 2 else p_[i % 2])
                    nm = names_all[(7 * i + 3) % len(names_all)]
                    if nm not in bubble_map.setdefault(v_, []):
                        bubble_map[v_].append(nm)
            bmap = _dd(list, {v_: list(l_) for v_, l_ in bubble_map.items()})
            for o2c_ in (A.multimap(A.get_ori2cov_onevertex, 1, all_paths, A.seqPaths, bmap, A.seqLimitsPerSeq, debug, chunksize=10) if all_paths else []):
                bubble_oris.append(sorted(o for o, cov in o2c_.items() if cov >= 1))      # before the "best origin" fallback and the _Nsplit_ cut

    edges = A.DBG._edges.astype(np.uint32)
    res = dict(
        joined=[(list(map(int, p_)), s_, list(o_)) for p_, s_, o_ in zip(joined, joined_seqs, joined_oris)],
        bubble_map=[[int(v_), list(l_)] for v_, l_ in sorted(bubble_map.items())],
        bubble_oris=bubble_oris,
        orient_counts=orient,
        orient_sets=orient_sets,
        k=ksize,
        names=list(A.seqDict),
        n_inst=int(sum(len(s) - ksize + 1 for s in A.seqDict.values() if len(s) > ksize)),
        vertex2coords=np.asarray(A.vertex2coords, dtype=np.uint32),
        edges=edges,
        seqPaths={n: np.asarray(sp, dtype=np.uint32) for n, sp in A.seqPaths.items()},
        seqLimits=sorted(int(v) for v in A.seqLimits),
        ncomp=len(comps),
        nbps=nbps_out,
    )
    return res


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    import digests
    r = run_reference(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 301,
                      threads=int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    import json
    print(json.dumps(digests.summarize(r), indent=1))
