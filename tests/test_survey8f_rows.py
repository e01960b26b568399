"""SURVEY 8(f) rows on the device.  Row 4: bubble candidates (NBPs sharing first and last vertex, lib/Assembler.py:728-734).
Row 1, second half: per raw NBP, the number of input sequences that contain it in their own orientation
(reference get_seqPaths_NBPs / nv2rightOrientation, lib/Assembler.py:1121-1139, :448-453).  The golden counts were
produced by the reference's own classmethod (oracle/refrun.py); -1 marks NBPs for which the reference defines no pair."""
import pytest

import dbg_oracle
import helpers
from superpang_b200 import _capi
from superpang_b200.assembler import Assembler

CASES = [c for c in helpers.load_crafted() if not c.get("degenerate")]


def _golden_map(case):
    r = case["result"]
    return {tuple(n[0]): c for n, c in zip(r["nbps"], r["orient_counts"]) if c >= 0}


def test_golden_defines_counts():
    assert sum(len(_golden_map(c)) for c in CASES) > 300


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['name']}-k{c['k']}")
def test_oracle_counts_match_reference(case):
    r = case["result"]
    seqs = [s for _, s in case["records"]]
    got = dbg_oracle.nbp_orientation_counts([(tuple(n[0]), n[1]) for n in r["nbps"]], seqs)
    for n, want, g in zip(r["nbps"], r["orient_counts"], got):
        if want >= 0:
            assert g == want, (case["name"], n[0][:4], want, g)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['name']}-k{c['k']}")
def test_emu_counts_match_reference(emu_lib, case):
    A = Assembler(dict(map(tuple, case["records"])), case["k"], engine=_capi.Engine(lib=emu_lib))
    got = A.nbp_orientation_counts()
    want = _golden_map(case)
    assert set(want) <= set(got)
    assert {k: got[k] for k in want} == want


def test_emu_counts_match_oracle_synthetic(emu_lib):
    for c in helpers.load_json("synthetic.json")["cases"][:3]:
        sd = helpers.synth_seqdict(c["n_genomes"], c["length"], c["ani"], c["seed"], c["mode"])
        A = Assembler(dict(sd), c["k"], engine=_capi.Engine(lib=emu_lib))
        raw = A.raw_nbps()
        want = dbg_oracle.nbp_orientation_counts(raw, list(sd.values()))
        got = A.nbp_orientation_counts()
        assert [got[r[0]] for r in raw] == want


@pytest.mark.gpu
def test_gpu_counts_match_reference_and_oracle():
    for case in CASES:
        A = Assembler(dict(map(tuple, case["records"])), case["k"], device=0)
        got = A.nbp_orientation_counts()
        want = _golden_map(case)
        assert {k: got[k] for k in want} == want, case["name"]
    for c in helpers.load_json("synthetic.json")["cases"]:
        sd = helpers.synth_seqdict(c["n_genomes"], c["length"], c["ani"], c["seed"], c["mode"])
        A = Assembler(dict(sd), c["k"], device=0)
        raw = A.raw_nbps()
        got = A.nbp_orientation_counts()
        assert [got[r[0]] for r in raw] == dbg_oracle.nbp_orientation_counts(raw, list(sd.values()))


def _norm(groups):
    return sorted(sorted(g) for g in groups)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['name']}-k{c['k']}")
def test_emu_bubble_candidates_match_reference_nbps(emu_lib, case):
    """grouping of the reference's own NBP list (golden) by (first, last) vertex."""
    want = dbg_oracle.bubble_candidates([(tuple(n[0]),) for n in case["result"]["nbps"]])
    A = Assembler(dict(map(tuple, case["records"])), case["k"], engine=_capi.Engine(lib=emu_lib))
    assert _norm(A.bubble_candidates()) == _norm(want)


def test_emu_bubble_candidates_synthetic(emu_lib):
    n_groups = 0
    for c in helpers.load_json("synthetic.json")["cases"][:3]:
        sd = helpers.synth_seqdict(c["n_genomes"], c["length"], c["ani"], c["seed"], c["mode"])
        A = Assembler(dict(sd), c["k"], engine=_capi.Engine(lib=emu_lib))
        want = dbg_oracle.bubble_candidates(A.raw_nbps())
        assert _norm(A.bubble_candidates()) == _norm(want)
        n_groups += len(want)
    assert n_groups > 10                                 # SNP bubbles: the case the reference's flattening is for


@pytest.mark.gpu
def test_gpu_bubble_candidates():
    for case in CASES:
        want = dbg_oracle.bubble_candidates([(tuple(n[0]),) for n in case["result"]["nbps"]])
        A = Assembler(dict(map(tuple, case["records"])), case["k"], device=0)
        assert _norm(A.bubble_candidates()) == _norm(want), case["name"]


def _golden_sets(case):
    """{name: set of vertex tuples} from the reference's own get_seqPaths_NBPs (only NBPs with a defined pair)."""
    r = case["result"]
    return {name: {tuple(r["nbps"][nv][0]) for nv in nvs} for name, nvs in r.get("orient_sets", {}).items()}


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['name']}-k{c['k']}")
def test_oracle_seq_sets_match_reference(case):
    r = case["result"]
    defined = {i for i, c in enumerate(r["orient_counts"]) if c >= 0}
    got = dbg_oracle.seq_paths_nbps([(tuple(n[0]), n[1]) for n in r["nbps"]], dict(map(tuple, case["records"])))
    got = {name: sorted(i for i in nvs if i in defined) for name, nvs in got.items()}
    got = {n: v for n, v in got.items() if v}
    assert got == {n: sorted(v) for n, v in r.get("orient_sets", {}).items()}


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['name']}-k{c['k']}")
def test_emu_seq_sets_match_reference(emu_lib, case):
    r = case["result"]
    defined = {tuple(n[0]) for n, c in zip(r["nbps"], r["orient_counts"]) if c >= 0}
    A = Assembler(dict(map(tuple, case["records"])), case["k"], engine=_capi.Engine(lib=emu_lib))
    got = {n: {v for v in s if v in defined} for n, s in A.seq_paths_nbps().items()}
    got = {n: s for n, s in got.items() if s}
    assert got == _golden_sets(case)


@pytest.mark.gpu
def test_gpu_seq_sets_match_reference():
    for case in CASES:
        r = case["result"]
        defined = {tuple(n[0]) for n, c in zip(r["nbps"], r["orient_counts"]) if c >= 0}
        A = Assembler(dict(map(tuple, case["records"])), case["k"], device=0)
        got = {n: {v for v in s if v in defined} for n, s in A.seq_paths_nbps().items()}
        assert {n: s for n, s in got.items() if s} == _golden_sets(case), case["name"]


# ---- row 2 (first half): reconstruct_sequence for arbitrary paths ---------------------------------------------------
def _check_paths(A, case):
    r = case["result"]
    paths = [n[0] for n in r["nbps"]] + [j[0] for j in r.get("joined", [])]
    want = [n[1] for n in r["nbps"]] + [j[1] for j in r.get("joined", [])]
    assert A.reconstruct_sequences(paths) == want
    # row 2, second half: origins of the same paths (raw NBPs: goldens of the NBP collection; joined pairs: the
    # reference's own get_ori2cov_onevertex + goodOris, oracle/refrun.py)
    want_o = [tuple(n[2]) for n in r["nbps"]] + [tuple(j[2]) for j in r.get("joined", [])]
    assert A.path_origins(paths) == want_o
    # the same joined paths handed over as lists of RAW NBP IDS (expanded on the device, spb_path_*_ex kind 1)
    voff, verts, _, _ = A._eng.nbps()
    ids = {tuple(verts[int(voff[i]):int(voff[i + 1])].tolist()): i for i in range(len(voff) - 1)}
    id_paths, vert_paths = [[ids[tuple(n[0])]] for n in r["nbps"][:5]], [n[0] for n in r["nbps"][:5]]
    for j in r.get("joined", []):
        p = tuple(j[0])
        for cut in range(2, len(p)):
            if p[:cut] in ids and p[cut - 1:] in ids:
                id_paths.append([ids[p[:cut]], ids[p[cut - 1:]]]); vert_paths.append(list(p))
                break
    if id_paths:
        assert A._eng.path_sequences(id_paths, case["k"], nbp_ids=True) == A._eng.path_sequences(vert_paths, case["k"])
        assert A._eng.path_origins(id_paths, nbp_ids=True) == A._eng.path_origins(vert_paths)
    # origins with a non-empty bubble_vertex2origins: the reference's own get_ori2cov_onevertex (cov >= 1) on every raw NBP
    # and joined path, with the made-up bubble map of oracle/refrun.py
    if r.get("bubble_oris"):
        index = {n: i for i, n in enumerate(A.ref2name)}
        bmap = {int(v): [index[n] for n in l] for v, l in r["bubble_map"]}
        got = A._eng.path_origins(paths, bubble_vertex2origins=bmap)
        assert [sorted(A.ref2name[i] for i in g) for g in got] == [sorted(w) for w in r["bubble_oris"]], case["name"]
    return len(r.get("joined", []))


def test_emu_path_sequences_match_reference(emu_lib):
    """Every golden NBP (incl. broken cycles, which step over the cut edge) and the joined NBP pairs whose sequences
    the reference's own reconstruct_sequence produced (oracle/refrun.py)."""
    n_joined = 0
    for case in CASES:
        A = Assembler(dict(map(tuple, case["records"])), case["k"], engine=_capi.Engine(lib=emu_lib))
        n_joined += _check_paths(A, case)
    assert n_joined > 50


def test_emu_path_sequences_reject_bad_paths(emu_lib):
    case = next(c for c in CASES if c["name"] == "snp_bubble")
    A = Assembler(dict(map(tuple, case["records"])), case["k"], engine=_capi.Engine(lib=emu_lib))
    nv = A.vertex2coords.shape[0]
    for bad in ([0], [0, nv // 2 + 7, 3], [0, 0], [nv + 5, 1]):
        with pytest.raises(_capi.SpbError):
            A.reconstruct_sequences([bad])
    A.compact()
    n = A.counts["n_nbp"]
    for bad in ([n + 3], [0, 0]):                                             # unknown NBP id; NBPs that do not share their junction vertex
        with pytest.raises(_capi.SpbError):
            A._eng.path_sequences([bad], case["k"], nbp_ids=True)
    assert A.reconstruct_sequences([]) == []
    assert len(A.reconstruct_sequences([[0, 1, 2]])[0]) == case["k"] + 2      # still usable after the errors


@pytest.mark.gpu
def test_gpu_path_sequences_match_reference():
    for case in CASES:
        _check_paths(Assembler(dict(map(tuple, case["records"])), case["k"], device=0), case)
