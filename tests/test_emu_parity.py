"""CPU tier: the CUDA kernels' LOGIC, executed by the host emulator build (tests/emu, -DSPB_EMU),
against the reference goldens and the oracle.  The emulator is test infrastructure; the product
never loads it.  The `-m gpu` tier runs the same checks through the real sm_100a library."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import dbg_oracle
import digests
import helpers
from superpang_b200 import _capi
from superpang_b200.assembler import Assembler


def run_emu(emu_lib, seqdict, k):
    return Assembler(dict(seqdict), k, engine=_capi.Engine(lib=emu_lib))


@pytest.mark.parametrize("case", helpers.load_crafted(), ids=lambda c: f"{c['name']}-k{c['k']}")
def test_emu_matches_reference_golden(emu_lib, case):
    if case.get("degenerate"):                      # the reference asserts here: the product must refuse
        with pytest.raises(_capi.DegenerateInputError):
            run_emu(emu_lib, dict(map(tuple, case["records"])), case["k"]).to_result()
        return
    A = run_emu(emu_lib, dict(map(tuple, case["records"])), case["k"])
    got = A.to_result()
    digests.assert_same(helpers.golden_to_result(case), got)
    helpers.assert_summary(digests.summarize(got), case["summary"])


@pytest.mark.parametrize("case", helpers.load_json("synthetic.json")["cases"],
                         ids=lambda c: f"{c['mode']}-seed{c['seed']}-k{c['k']}")
def test_emu_matches_reference_synthetic(emu_lib, case):
    sd = helpers.synth_seqdict(case["n_genomes"], case["length"], case["ani"], case["seed"], case["mode"])
    helpers.assert_summary(digests.summarize(run_emu(emu_lib, sd, case["k"]).to_result()), case["summary"])


def test_emu_matches_reference_real_slice(emu_lib):
    want = helpers.load_json("real_slice.json")
    got = digests.summarize(run_emu(emu_lib, helpers.real_slice_records(), want["k"]).to_result())
    helpers.assert_summary(got, want["summary"])


_COMP = str.maketrans("ACGT", "TGCA")


@st.composite
def genome_sets(draw):
    """Small random pangenomes built from shared blocks, inversions, duplications and point edits:
    planted branches, tips, loops, shared ends and repeats."""
    rng = np.random.default_rng(draw(st.integers(0, 2**32 - 1)))
    nblocks = draw(st.integers(2, 6))
    blocks = ["".join("ACGT"[c] for c in rng.integers(0, 4, int(rng.integers(20, 120)))) for _ in range(nblocks)]
    nseq = draw(st.integers(1, 5))
    seqs = {}
    for i in range(nseq):
        parts = []
        for _ in range(int(rng.integers(1, 6))):
            b = blocks[int(rng.integers(0, nblocks))]
            if rng.random() < 0.3:
                b = b.translate(_COMP)[::-1]
            if rng.random() < 0.3:
                j = int(rng.integers(0, len(b)))
                b = b[:j] + "ACGT"[int(rng.integers(0, 4))] + b[j + 1:]
            parts.append(b)
        s = "".join(parts)
        if rng.random() < 0.15:
            s = s + s[:int(rng.integers(1, 40))]
        seqs[f"s{i}"] = s
    return seqs


@settings(max_examples=int(os.environ.get("SPB_HYP_EXAMPLES", "120")), deadline=None, suppress_health_check=list(HealthCheck))
@given(genome_sets(), st.sampled_from([9, 15, 21, 31]))
def test_emu_matches_oracle_random(emu_lib, seqs, k):
    try:
        want = dbg_oracle.run_oracle(seqs, k)
    except (AssertionError, IndexError):
        # degenerate input: the reference itself asserts; the product must refuse, not differ
        with pytest.raises(_capi.SpbError):
            run_emu(emu_lib, seqs, k).to_result()
        return
    if not want["vertex2coords"].shape[0]:
        return
    A = run_emu(emu_lib, seqs, k)
    digests.assert_same(want, A.to_result())
    raw = A.raw_nbps()                                   # orientation counts against the oracle
    got = A.nbp_orientation_counts()
    assert [got[r[0]] for r in raw] == dbg_oracle.nbp_orientation_counts(raw, list(seqs.values()))
    want_sets = {n: {raw[i][0] for i in ids} for n, ids in dbg_oracle.seq_paths_nbps(raw, seqs).items()}
    assert A.seq_paths_nbps() == want_sets
    assert sorted(sorted(g) for g in A.bubble_candidates()) == sorted(sorted(g) for g in dbg_oracle.bubble_candidates(raw))
    assert A.reconstruct_sequences([r[0] for r in raw]) == [r[1] for r in raw]       # row 2: the path getters on the raw NBPs
    assert A.path_origins([r[0] for r in raw]) == [r[2] for r in raw]


@pytest.mark.parametrize("k,env", [(1001, {}), (2001, {}), (1001, {"SPB_DEDUP": "smem", "SPB_SMEM_AVG": "500"}), (75, {"SPB_DEDUP": "smem"}),
                                   (1001, {"SPB_DEDUP": "part", "SPB_PART_AVG": "500"}), (75, {"SPB_DEDUP": "part", "SPB_PART_AVG": "100"})])
def test_emu_large_k_matches_oracle(emu_lib, monkeypatch, k, env):
    """k beyond the specialised widths (101 / 301 / 501): the generic limb loops, the rolling hashes with a long window."""
    for key, v in env.items():
        monkeypatch.setenv(key, v)
    sd = helpers.synth_seqdict(3, 12000 if k > 500 else 4000, 0.99, 7, "block")
    want = dbg_oracle.run_oracle(sd, k)
    digests.assert_same(want, run_emu(emu_lib, sd, k).to_result())


@pytest.mark.parametrize("k", [3, 9, 15, 17, 19])
@pytest.mark.parametrize("env", [{}, {"SPB_DEDUP": "part", "SPB_PART_AVG": "100"}, {"SPB_DEDUP": "smem", "SPB_SMEM_AVG": "100"},
                                 {"SPB_DEDUP": "links", "SPB_BUCKET_BITS": "3"}, {"SPB_DEDUP": "staged", "SPB_BUCKET_BITS": "3"}],
                         ids=lambda e: ",".join(f"{k_[4:].lower()}={v}" for k_, v in e.items()) or "default")
def test_emu_tiny_k_matches_oracle(emu_lib, monkeypatch, k, env):
    """k below and just above the 16-base windows of the rolling-hash kernels: those paths must hand tiny k over to the
    kernels that hash all limbs (k < 17) instead of reading in front of the k-mer, and agree with the oracle either way."""
    for key, v in env.items():
        monkeypatch.setenv(key, v)
    rng = np.random.default_rng(100 + k)
    anc = rng.integers(0, 4, 3000)
    sd = {}
    for g in range(3):
        x = anc.copy()
        m = rng.random(x.size) < 0.03
        x[m] = (x[m] + rng.integers(1, 4, int(m.sum()))) & 3
        sd[f"g{g}|c"] = "".join("ACGT"[v] for v in x)
    try:
        want = dbg_oracle.run_oracle(sd, k)
    except AssertionError:                              # orientation-ambiguous at this k: the new path must refuse as well
        with pytest.raises(_capi.DegenerateInputError):
            run_emu(emu_lib, sd, k).to_result()
        return
    digests.assert_same(want, run_emu(emu_lib, sd, k).to_result())


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_emu_optimistic_links_fuzz(emu_lib, monkeypatch, seed):
    """Optimistic links + verification + repair against the oracle on random genome sets (SNPs, reverse-complemented and
    rotated copies), with random subsets of the stored hash bits masked away (so that distinct k-mers collide and links
    fail), random bucket sizes and thread orders."""
    import random
    rnd = random.Random(seed)
    checked = 0
    for _ in range(14):
        k = rnd.choice([17, 21, 31, 33, 45])
        n, L, p = rnd.randint(2, 5), rnd.randint(200, 2000), rnd.choice([0.0, 0.01, 0.03, 0.1])
        rng = np.random.default_rng(rnd.randint(0, 1 << 30))
        anc = rng.integers(0, 4, L)
        sd = {}
        for g in range(n):
            x = anc.copy()
            m = rng.random(L) < p
            x[m] = (x[m] + rng.integers(1, 4, int(m.sum()))) & 3
            if rnd.random() < 0.3:
                x = (3 - x)[::-1]
            if rnd.random() < 0.3:
                a = rnd.randint(0, L - 50)
                x = np.concatenate([x[a:], x[:a]])
            sd[f"g{g}"] = "".join("ACGT"[v] for v in x)
        for key, v in dict(SPB_DEDUP="part", SPB_PART_AVG=str(rnd.choice([20, 60, 300])),
                           SPB_D_HMASK=rnd.choice(["FFFFFFFF", "FFF00000", "FF000000", "FFFF0000", "F8000000"]),
                           SPB_EMU_ORDER=rnd.choice(["", "reverse", "shuffle"])).items():
            monkeypatch.setenv(key, v)
        try:
            want = dbg_oracle.run_oracle(sd, k)
        except AssertionError:
            with pytest.raises(_capi.DegenerateInputError):
                run_emu(emu_lib, sd, k).to_result()
            continue
        digests.assert_same(want, run_emu(emu_lib, sd, k).to_result())
        checked += 1
    assert checked >= 10


def test_rejects_bad_input(emu_lib):
    with pytest.raises(_capi.SpbError):
        run_emu(emu_lib, {"a": "ACGT" * 20}, 30)           # even k
    with pytest.raises(_capi.SpbError):
        run_emu(emu_lib, {"a": "ACGTRYACGT" * 10}, 9)       # non-ACGT reaches the device
    with pytest.raises(NotImplementedError):
        Assembler({"a": "ACGT" * 20}, 9, gap_size=2, engine=_capi.Engine(lib=emu_lib))


VARIANTS = [
    {"SPB_EMU_ORDER": "reverse"},                                   # larger positions reach the table first (displacements)
    {"SPB_EMU_ORDER": "shuffle"},
    {"SPB_DEDUP": "staged", "SPB_BUCKET_BITS": "3"},                # partitioned dedup, 8 buckets
    {"SPB_DEDUP": "staged", "SPB_BUCKET_BITS": "5", "SPB_EMU_ORDER": "reverse"},
    {"SPB_ANCHOR": "1", "SPB_CHUNK_BITS": "8"},                      # anchor + extend forced, 256-position chunks
    {"SPB_ANCHOR": "1", "SPB_CHUNK_BITS": "9", "SPB_EMU_ORDER": "reverse"},
    {"SPB_CHUNK_BITS": "10"},                                        # adaptive decision after the second chunk
    {"SPB_ANCHOR": "0", "SPB_CHUNK_BITS": "8", "SPB_EMU_ORDER": "shuffle"},
    {"SPB_DEDUP": "links", "SPB_BUCKET_BITS": "3"},                  # partitioned dedup, exceptions as links
    {"SPB_DEDUP": "links", "SPB_BUCKET_BITS": "0", "SPB_EMU_ORDER": "reverse"},   # displacement chains
    {"SPB_DEDUP": "links", "SPB_BUCKET_BITS": "5", "SPB_EMU_ORDER": "shuffle"},
    {"SPB_DEDUP": "links", "SPB_BUCKET_BITS": "4", "SPB_PARTITION_FUSED": "1", "SPB_EMU_ORDER": "shuffle"},   # block-phased partition kernel
    {"SPB_DEDUP": "smem"},                                           # shared-memory tables, one sub-bucket
    {"SPB_DEDUP": "smem", "SPB_SMEM_AVG": "300", "SPB_EMU_ORDER": "reverse"},    # several sub-buckets; larger index installs first
    {"SPB_DEDUP": "smem", "SPB_SMEM_AVG": "40", "SPB_EMU_ORDER": "shuffle"},
    {"SPB_SWITCH_MIN": "1", "SPB_CHUNK_BITS": "8", "SPB_SMEM_AVG": "100"},   # adaptive default: small-table probe, then one of the three schemes
    {"SPB_SWITCH_MIN": "1", "SPB_CHUNK_BITS": "9", "SPB_NO_SMEM_TABLES": "1", "SPB_EMU_ORDER": "reverse"},
    {"SPB_DEDUP": "smem", "SPB_SMEM_LIMIT": "10"},                   # a sub-bucket beyond the table -> L2-table path instead
    {"SPB_DEDUP": "smem", "SPB_SMEM_AVG": "1"},                      # absurd split: few hash bits left -> falls back or clusters
    {"SPB_DEDUP": "part"},                                           # hand-written two-level partition, one bucket
    {"SPB_DEDUP": "part", "SPB_PART_AVG": "300", "SPB_EMU_ORDER": "reverse"},     # several buckets, both levels; larger positions first
    {"SPB_DEDUP": "part", "SPB_PART_AVG": "40", "SPB_EMU_ORDER": "shuffle"},
    {"SPB_DEDUP": "part", "SPB_PART_AVG": "40", "SPB_PART_L1": "0"},              # level 2 only
    {"SPB_DEDUP": "part", "SPB_PART_AVG": "2"},                      # buckets overflow their regions -> sort-based path instead
    {"SPB_DEDUP": "part", "SPB_D_EXACT": "1", "SPB_PART_AVG": "300"},                 # inline exact verification instead of optimistic links
    # optimistic links with stored hash bits masked away: distinct k-mers collide, their links fail the verification pass and
    # are re-homed by the repair kernels (normally a handful per 10^9 records)
    {"SPB_DEDUP": "part", "SPB_D_HMASK": "FFF00000", "SPB_PART_AVG": "300"},
    {"SPB_DEDUP": "part", "SPB_D_HMASK": "FFF00000", "SPB_PART_AVG": "300", "SPB_EMU_ORDER": "reverse"},   # take-over chains through wrong links
    {"SPB_DEDUP": "part", "SPB_D_HMASK": "FFC00000", "SPB_PART_AVG": "100", "SPB_EMU_ORDER": "shuffle"},
    {"SPB_DEDUP": "part", "SPB_D_HMASK": "F0000000", "SPB_PART_AVG": "300"},          # more failures than the repair list holds -> exact path
    {"SPB_DEDUP": "part", "SPB_D_HMASK": "0", "SPB_PART_AVG": "40"},                  # every record of a bucket collides
    {"SPB_DEDUP": "part", "SPB_D_HMASK": "80000000", "SPB_D_EXACT": "1", "SPB_PART_AVG": "100"},
    {"SPB_SWITCH_MIN": "1", "SPB_CHUNK_BITS": "8", "SPB_PART_AVG": "100", "SPB_EMU_ORDER": "reverse"},   # adaptive default reaching the partition
]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k[4:].lower()}={v}" for k, v in e.items()))
def test_emu_variants_match_reference_golden(emu_lib, monkeypatch, env):
    """The dedup paths that depend on GPU scheduling (displaced first positions, exact re-probe) and the
    partitioned dedup, forced in the emulator through thread order / path selection."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for case in helpers.load_crafted():
        if case.get("degenerate"):
            continue
        got = run_emu(emu_lib, dict(map(tuple, case["records"])), case["k"]).to_result()
        digests.assert_same(helpers.golden_to_result(case), got)
    for case in helpers.load_json("synthetic.json")["cases"][:3]:
        sd = helpers.synth_seqdict(case["n_genomes"], case["length"], case["ani"], case["seed"], case["mode"])
        helpers.assert_summary(digests.summarize(run_emu(emu_lib, sd, case["k"]).to_result()), case["summary"])


def test_all_sequences_shorter_than_k(emu_lib):
    """Every sequence has len <= k: the reference builds an empty graph and yields no contigs (`lib/Assembler.py:150-153`);
    the new path must do the same instead of failing in the compaction."""
    from superpang_b200 import _capi
    from superpang_b200.assembler import Assembler
    A = Assembler({"a": "ACGTACGTAC", "b": "ACG"}, 31, engine=_capi.Engine(lib=emu_lib))
    assert A.counts["n_inst"] == 0 and A.counts["n_vertices"] == 0
    c = A.compact()
    assert c["n_nbp"] == 0 and c["n_comp"] == 0
    assert A.collect_nbps() == [] and A.vertex2coords.shape == (0, 2) and A.edges.shape == (0, 2) and A.seqPaths == {}


def test_distributed_build_validates_world_size(emu_lib):
    """world size 1 falls back to the single-GPU build; a non-power-of-two size is refused before any collective."""
    import torch.distributed as dist
    from superpang_b200 import _capi

    class _FakeDist:
        def __init__(self, n):
            self.n = n
    eng = _capi.Engine(lib=emu_lib)
    import numpy as np
    seq = np.frombuffer(b"ACGTTGCATGGACCATTGACGGTACCATGACAGTTGACCATGAAACCCGTGTTAGC", dtype=np.uint8)
    eng.load(seq, np.array([0, len(seq)], dtype=np.uint64), 31)
    real = (dist.get_world_size, dist.get_rank)
    try:
        dist.get_world_size, dist.get_rank = (lambda group=None: 3), (lambda group=None: 0)
        with pytest.raises(ValueError):
            eng.build_dbg_distributed()
        dist.get_world_size = lambda group=None: 1
        assert eng.build_dbg_distributed()["n_inst"] == len(seq) - 30
    finally:
        dist.get_world_size, dist.get_rank = real
