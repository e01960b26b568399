"""
TEST INFRASTRUCTURE ONLY -- canonical digests of a DBG/NBP result (definitions: SURVEY.md §8(c)).

A "result" is the dict produced by `oracle/refrun.py:run_reference`, `oracle/dbg_oracle.py:run_oracle`
and `superpang_b200.assembler` (`.to_result()`), so the three can be compared field by field or by
digest.  NBPs are compared as a multiset after canonical ordering (the parity contract of
BASELINE.json: identical NBP sequence set and identical NBP->origin assignments).
"""
import hashlib

import numpy as np


def _sha(b):
    return hashlib.sha256(b).hexdigest()


def d_seq(nbps):
    return _sha("\n".join(sorted(s for _, s, *_ in nbps)).encode())


def d_ori(nbps):
    return _sha("\n".join(sorted(f"{s}\t{','.join(sorted(o))}" for _, s, o, *_ in nbps)).encode())


def d_nbpv(nbps):
    return _sha("\n".join(sorted(",".join(map(str, v)) for v, *_ in nbps)).encode())


def d_seqpaths(res):
    h = hashlib.sha256()
    for n in res["names"]:
        sp = res["seqPaths"].get(n)
        h.update(n.encode() + b"\0")
        if sp is not None:
            h.update(np.ascontiguousarray(sp, dtype=np.uint32).tobytes())
        h.update(b"\1")
    return h.hexdigest()


def summarize(res):
    nbps = res["nbps"]
    lens = [len(s) for _, s, *_ in nbps]
    return dict(
        k=res["k"],
        n_seq=len(res["names"]),
        n_inst=res["n_inst"],
        n_vertices=int(res["vertex2coords"].shape[0]),
        n_edges=int(res["edges"].shape[0]),
        n_comp=res["ncomp"],
        n_nbp=len(nbps),
        nbp_bases=int(sum(lens)),
        nbp_len_min=int(min(lens)) if lens else 0,
        nbp_len_max=int(max(lens)) if lens else 0,
        n_nbp_2v=sum(1 for v, *_ in nbps if len(v) == 2),
        n_cycle_nbp=sum(1 for *_, cyc in nbps if cyc),
        sha_vertex2coords=_sha(np.ascontiguousarray(res["vertex2coords"], dtype=np.uint32).tobytes()),
        sha_edges=_sha(np.ascontiguousarray(res["edges"], dtype=np.uint32).tobytes()),
        sha_seqlimits=_sha(np.asarray(res["seqLimits"], dtype=np.uint32).tobytes()),
        sha_seqpaths=d_seqpaths(res),
        D_nbpv=d_nbpv(nbps),
        D_seq=d_seq(nbps),
        D_ori=d_ori(nbps),
    )


def assert_same(a, b, what=("a", "b")):
    """Field-by-field comparison with readable failure messages."""
    assert a["k"] == b["k"]
    assert a["names"] == b["names"], "sequence names differ"
    assert a["n_inst"] == b["n_inst"], (a["n_inst"], b["n_inst"])
    va, vb = np.asarray(a["vertex2coords"]), np.asarray(b["vertex2coords"])
    assert va.shape == vb.shape, f"vertex count {va.shape} vs {vb.shape}"
    assert np.array_equal(va, vb), "vertex2coords differ"
    ea, eb = np.asarray(a["edges"]), np.asarray(b["edges"])
    assert ea.shape == eb.shape, f"edge count {ea.shape} vs {eb.shape}"
    assert np.array_equal(ea, eb), "edges differ"
    assert list(a["seqLimits"]) == list(b["seqLimits"]), "seqLimits differ"
    assert set(a["seqPaths"]) == set(b["seqPaths"]), "seqPaths keys differ"
    for n in a["seqPaths"]:
        assert np.array_equal(a["seqPaths"][n], b["seqPaths"][n]), f"seqPath {n} differs"
    assert a["ncomp"] == b["ncomp"], (a["ncomp"], b["ncomp"])
    na = sorted((v, s, tuple(o), bool(c)) for v, s, o, _, c in a["nbps"])
    nb = sorted((v, s, tuple(o), bool(c)) for v, s, o, _, c in b["nbps"])
    assert len(na) == len(nb), f"NBP count {len(na)} vs {len(nb)}"
    for x, y in zip(na, nb):
        assert x[0] == y[0], f"NBP vertices differ: {x[0][:8]}.. vs {y[0][:8]}.."
        assert x[1] == y[1], f"NBP sequence differs for {x[0][:8]}"
        assert x[2] == y[2], f"NBP origins differ for {x[0][:8]}: {x[2]} vs {y[2]}"
        assert x[3] == y[3], f"NBP cycle flag differs for {x[0][:8]}"
    # component partition (labels may be numbered differently: compare the grouping of NBPs)
    ga, gb = {}, {}
    for v, _, _, c, _ in a["nbps"]:
        ga.setdefault(c, set()).add(v)
    for v, _, _, c, _ in b["nbps"]:
        gb.setdefault(c, set()).add(v)
    assert sorted(map(sorted, ga.values())) == sorted(map(sorted, gb.values())), "component grouping differs"
