"""GPU tier (`-m gpu`): the product path -- libsuperpang_b200.so (sm_100a) through the C-ABI -- against the
reference goldens, the oracle on seeded inputs, the KAT digests of the bundled genomes, and
size-independent properties at benchmark scale.  Bit-exact: integer / byte work only."""
import os

import numpy as np
import pytest

import dbg_oracle
import digests
import helpers
from superpang_b200 import _capi, synth
from superpang_b200.assembler import Assembler, read_fasta

pytestmark = pytest.mark.gpu


def run_gpu(seqdict, k):
    A = Assembler(seqdict, k, device=0)
    assert A._eng.kind == "cuda sm_100a"
    return A


@pytest.mark.parametrize("case", helpers.load_crafted(), ids=lambda c: f"{c['name']}-k{c['k']}")
def test_gpu_matches_reference_golden(case):
    if case.get("degenerate"):                      # the reference asserts here: the product must refuse
        with pytest.raises(_capi.DegenerateInputError):
            run_gpu(dict(map(tuple, case["records"])), case["k"]).to_result()
        return
    got = run_gpu(dict(map(tuple, case["records"])), case["k"]).to_result()
    digests.assert_same(helpers.golden_to_result(case), got)
    helpers.assert_summary(digests.summarize(got), case["summary"])


@pytest.mark.parametrize("case", helpers.load_json("synthetic.json")["cases"],
                         ids=lambda c: f"{c['mode']}-seed{c['seed']}-k{c['k']}")
def test_gpu_matches_reference_synthetic(case):
    sd = helpers.synth_seqdict(case["n_genomes"], case["length"], case["ani"], case["seed"], case["mode"])
    helpers.assert_summary(digests.summarize(run_gpu(sd, case["k"]).to_result()), case["summary"])


def test_gpu_matches_reference_real_slice():
    want = helpers.load_json("real_slice.json")
    helpers.assert_summary(digests.summarize(run_gpu(helpers.real_slice_records(), want["k"]).to_result()), want["summary"])


@pytest.mark.parametrize("mode,k", [("block", 301), ("snp", 301), ("block", 101), ("block", 501)])
def test_gpu_matches_oracle_medium(mode, k):
    """~0.4 M instances: field-by-field against the oracle."""
    sd = helpers.synth_seqdict(6, 70000, 0.96, 31, mode)
    digests.assert_same(dbg_oracle.run_oracle(sd, k), run_gpu(sd, k).to_result())


@pytest.mark.parametrize("which", ["A", "B"])
def test_gpu_kat_bundled_genomes(which):
    """KAT digests of the reference's bundled genomes; data staged by build()."""
    fa = os.path.join(helpers.DATA, f"{which}.fa")
    if not os.path.exists(fa):
        pytest.skip("tests/_data not staged (reference test data absent at build time)")
    A = Assembler(fa, 301, device=0)
    helpers.assert_summary(digests.summarize(A.to_result()), helpers.load_json(f"kat_{which}.json"))


def _codes(s):
    lut = np.zeros(256, np.uint8)
    lut[list(b"ACGT")] = [0, 1, 2, 3]
    return lut[np.frombuffer(s.encode() if isinstance(s, str) else s, dtype=np.uint8)]


@pytest.mark.parametrize("mode", ["snp", "block"])
def test_gpu_properties_at_scale(mode):
    """BASELINE config 2 (10 genomes x 2 Mbp, k=301): properties that do not need the oracle."""
    k = 301
    names, seqs = synth.genome_set(10, 2_000_000, 0.97, seed=1, mode=mode)
    eng = _capi.Engine(device=0)
    from superpang_b200.assembler import concat_sequences
    buf, off = concat_sequences(seqs)
    eng.load(buf, off, k)
    c = eng.build_dbg()
    c = eng.compact()
    assert c["n_inst"] == synth.n_instances(seqs, k)
    sp = eng.instance_vertices()
    v2c = eng.vertex2coords()
    valid = sp != _capi.SPB_NONE
    assert int(valid.sum()) == c["n_inst"]
    # vertex ids = order of first appearance: first occurrence positions are strictly increasing
    lin = off[v2c[:, 0].astype(np.int64)].astype(np.int64) + v2c[:, 1].astype(np.int64)
    assert np.all(np.diff(lin) > 0)
    assert np.array_equal(sp[lin], np.arange(c["n_vertices"], dtype=np.uint32))
    assert int(sp[valid].max()) == c["n_vertices"] - 1
    # edges: exactly the unique unordered consecutive-instance pairs
    pair = valid[:-1] & valid[1:]
    a, b = sp[:-1][pair].astype(np.uint64), sp[1:][pair].astype(np.uint64)
    key = np.unique((np.minimum(a, b) << np.uint64(32)) | np.maximum(a, b))
    e = eng.edges().astype(np.uint64)
    assert np.array_equal((e[:, 0] << np.uint64(32)) | e[:, 1], key)
    # NBPs: every vertex is an init (end of its NBPs) or interior to exactly one NBP pair
    voff, verts, comp, cyc = eng.nbps()
    soff, sq = eng.nbp_seqs()
    assert np.all((soff[1:] - soff[:-1]) == (voff[1:] - voff[:-1]) + np.uint64(k - 1))
    L = (voff[1:] - voff[:-1]).astype(np.int64)
    interior = np.ones(verts.size, bool)
    interior[voff[:-1].astype(np.int64)] = False
    interior[voff[1:].astype(np.int64) - 1] = False
    cnt = np.bincount(verts[interior], minlength=c["n_vertices"])
    assert set(np.unique(cnt)) <= {0, 2}
    ends = np.zeros(c["n_vertices"], bool)
    ends[verts[~interior]] = True
    assert np.all(ends | (cnt == 2))
    # reverse-complement closure of the NBP sequence multiset (by hash)
    import hashlib
    comp_tab = bytes.maketrans(b"ACGT", b"TGCA")
    sb = sq.tobytes()
    hs, hr = [], []
    for i in range(len(L)):
        s = sb[int(soff[i]):int(soff[i + 1])]
        hs.append(hashlib.sha1(s).digest())
        hr.append(hashlib.sha1(s.translate(comp_tab)[::-1]).digest())
    assert sorted(hs) == sorted(hr)
    # a sample of NBPs re-k-merises to exactly its vertices (canonical k-mer of the vertex' first appearance)
    rng = np.random.default_rng(0)
    allseq = np.concatenate(seqs)
    for i in rng.choice(len(L), size=min(40, len(L)), replace=False):
        s = np.frombuffer(sb[int(soff[i]):int(soff[i + 1])], dtype=np.uint8)
        vs = verts[int(voff[i]):int(voff[i + 1])]
        for r in rng.choice(len(vs), size=min(25, len(vs)), replace=False):
            km = s[r:r + k].tobytes()
            p = int(lin[vs[r]])
            ref = allseq[p:p + k].tobytes()
            assert km == ref or km == ref.translate(comp_tab)[::-1]
    # origins are non-empty, sorted, in range
    ooff, oidx = eng.nbp_origins()
    assert np.all(ooff[1:] > ooff[:-1])
    assert int(oidx.max()) < len(seqs)


@pytest.mark.parametrize("env", [
    {"SPB_ANCHOR": "1", "SPB_CHUNK_BITS": "12"},       # anchor + extend forced with 4096-position chunks
    {"SPB_ANCHOR": "0", "SPB_CHUNK_BITS": "12"},
    {"SPB_DEDUP": "staged", "SPB_BUCKET_BITS": "4"},   # partitioned dedup, 16 buckets
    {"SPB_DEDUP": "links", "SPB_BUCKET_BITS": "4"},    # partitioned dedup, exceptions as links
    {"SPB_DEDUP": "smem"},                             # shared-memory tables
    {"SPB_DEDUP": "smem", "SPB_SMEM_AVG": "50"},       # ... many small sub-buckets
    {"SPB_DEDUP": "part"},                             # hand-written two-level partition + bulk-copied buckets
    {"SPB_DEDUP": "part", "SPB_PART_AVG": "50"},       # ... many small buckets, both levels
    {"SPB_DEDUP": "part", "SPB_PART_AVG": "200", "SPB_P1_SHAPE": "1", "SPB_P1_COPIES": "4", "SPB_SWITCH_MIN": "1"},   # ... the other CTA shapes of the fused kernel, region copies
    {"SPB_DEDUP": "part", "SPB_PART_AVG": "200", "SPB_P1_SHAPE": "2", "SPB_D_NBUF": "2", "SPB_D_EXACT": "1"},          # ... persistent double-buffered dedup kernel (inline verification)
    {"SPB_DEDUP": "part", "SPB_PART_AVG": "200", "SPB_D_EXACT": "1"},                            # ... inline warp-cooperative verification instead of optimistic links
    {"SPB_DEDUP": "part", "SPB_PART_AVG": "200", "SPB_D_HMASK": "FFF00000"},                     # ... optimistic links with provoked hash collisions: verification pass + repair kernels
], ids=lambda e: ",".join(f"{k[4:].lower()}={v}" for k, v in e.items()))
def test_gpu_dedup_variants(monkeypatch, env):
    """Every dedup strategy must give the reference's result (they differ only in how the table is reached)."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    for case in helpers.load_json("synthetic.json")["cases"]:
        sd = helpers.synth_seqdict(case["n_genomes"], case["length"], case["ani"], case["seed"], case["mode"])
        helpers.assert_summary(digests.summarize(run_gpu(sd, case["k"]).to_result()), case["summary"])
    want = helpers.load_json("real_slice.json")
    helpers.assert_summary(digests.summarize(run_gpu(helpers.real_slice_records(), want["k"]).to_result()), want["summary"])
    fa = os.path.join(helpers.DATA, "A.fa")
    if os.path.exists(fa):
        helpers.assert_summary(digests.summarize(Assembler(fa, 301, device=0).to_result()), helpers.load_json("kat_A.json"))


def test_gpu_multi_rank_parity():
    """>= 2 GPUs: hash-partitioned DBG build over NCCL, every rank must reproduce the reference goldens."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(helpers.ROOT, "tests", "mg_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multi-GPU parity OK" in r.stdout


@pytest.mark.parametrize("mode", ["snp", "block"])
def test_gpu_dedup_strategies_agree_at_scale(monkeypatch, mode):
    """BASELINE config 2 scale (20 M instances): every dedup strategy must produce byte-identical device results
    (vertex table, edge list, per-instance ids, NBP vertex lists / sequences / origins)."""
    import hashlib
    from superpang_b200.assembler import concat_sequences
    names, seqs = synth.genome_set(10, 2_000_000, 0.97, seed=1, mode=mode)
    buf, off = concat_sequences(seqs)

    def run(env):
        for k_ in ("SPB_DEDUP", "SPB_ANCHOR", "SPB_BUCKET_BITS", "SPB_CHUNK_BITS", "SPB_SMEM_AVG", "SPB_PART_AVG", "SPB_P1_SHAPE", "SPB_P1_COPIES", "SPB_D_NBUF"):
            monkeypatch.delenv(k_, raising=False)
        for k_, v in env.items():
            monkeypatch.setenv(k_, v)
        eng = _capi.Engine(device=0)
        eng.load(buf, off, 301)
        eng.build_dbg()
        eng.compact()
        h = hashlib.sha256()
        for arr in (eng.vertex2coords(), eng.edges(), eng.instance_vertices(), *eng.nbps(), *eng.nbp_seqs(), *eng.nbp_origins()):
            h.update(np.ascontiguousarray(arr).tobytes())
        eng.close()
        return h.hexdigest()

    ref = run({"SPB_DEDUP": "global", "SPB_ANCHOR": "0"})
    assert run({"SPB_DEDUP": "global", "SPB_ANCHOR": "1"}) == ref
    assert run({"SPB_DEDUP": "links"}) == ref
    assert run({"SPB_DEDUP": "smem"}) == ref
    assert run({"SPB_DEDUP": "smem", "SPB_SMEM_AVG": "2000"}) == ref
    assert run({"SPB_DEDUP": "part"}) == ref
    assert run({"SPB_DEDUP": "part", "SPB_PART_AVG": "1000"}) == ref
    assert run({"SPB_DEDUP": "staged"}) == ref
    assert run({}) == ref
