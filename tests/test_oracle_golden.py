"""CPU tier: the oracle restatement must reproduce what the UNMODIFIED reference produced
(golden vectors frozen by oracle/make_golden.py and the KAT digests of SURVEY.md §8(c))."""
import os

import numpy as np
import pytest

import dbg_oracle
import digests
import helpers


@pytest.mark.parametrize("case", helpers.load_crafted(), ids=lambda c: f"{c['name']}-k{c['k']}")
def test_oracle_reproduces_reference_crafted(case):
    if case.get("degenerate"):
        with pytest.raises(AssertionError):
            dbg_oracle.run_oracle(dict(map(tuple, case["records"])), case["k"])
        return
    got = dbg_oracle.run_oracle(dict(map(tuple, case["records"])), case["k"])
    digests.assert_same(helpers.golden_to_result(case), got)
    helpers.assert_summary(digests.summarize(got), case["summary"])


@pytest.mark.parametrize("case", helpers.load_json("synthetic.json")["cases"],
                         ids=lambda c: f"{c['mode']}-seed{c['seed']}-k{c['k']}")
def test_oracle_reproduces_reference_synthetic(case):
    sd = helpers.synth_seqdict(case["n_genomes"], case["length"], case["ani"], case["seed"], case["mode"])
    helpers.assert_summary(digests.summarize(dbg_oracle.run_oracle(sd, case["k"])), case["summary"])


def test_oracle_reproduces_reference_real_slice():
    want = helpers.load_json("real_slice.json")
    got = digests.summarize(dbg_oracle.run_oracle(helpers.real_slice_records(), want["k"]))
    helpers.assert_summary(got, want["summary"])


def test_compress_restatement_and_order():
    """compress() (Compressor.pyx:32-58) is order preserving, so max over raw k-mers == max over hashes."""
    rng = np.random.default_rng(5)
    for k in (31, 33, 101, 301):
        ks = ["".join("ACGT"[c] for c in rng.integers(0, 4, k)) for _ in range(60)]
        ks += [ks[0][:-1] + b for b in "ACGT"] + [b + ks[1][1:] for b in "ACGT"]
        hs = [dbg_oracle.compress(x) for x in ks]
        assert all(len(h) == k // 4 + k % 4 for h in hs)
        for a, ha in zip(ks, hs):
            for b, hb in zip(ks, hs):
                assert (a < b) == (ha < hb) and (a == b) == (ha == hb)


def test_compress_matches_reference_compiled_compressor():
    """Only where oracle/_ref (the reference's own Compressor.pyx, compiled) is present."""
    import glob
    import importlib.util
    so = glob.glob(os.path.join(helpers.ROOT, "oracle", "_ref", "superpang_lib", "Compressor.*.so"))
    if not so:
        pytest.skip("oracle/_ref not built (reference absent)")
    spec = importlib.util.spec_from_file_location("Compressor", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(6)
    for k in (31, 34, 101, 301, 503):
        comp = mod.Compressor(k)
        for _ in range(40):
            x = "".join("ACGT"[c] for c in rng.integers(0, 4, k)).encode()
            assert bytes(comp.compress(x)) == dbg_oracle.compress(x)


def test_compress_vertices_quirk():
    assert dbg_oracle.compress_vertices([1, 2, 3, 7]).tolist() == [[0, 3], [7, 7]]     # SURVEY §8a row 6 [measured]
    assert dbg_oracle.compress_vertices([2, 3, 9]).tolist() == [[2, 3], [9, 9]]
    assert dbg_oracle.compress_vertices([5, 0, 1]).tolist() == [[0, 1], [5, 5]]


@pytest.mark.slow
@pytest.mark.skipif(not os.environ.get("SPB_SLOW"), reason="set SPB_SLOW=1 (about a minute)")
def test_oracle_kat_A():
    fa = os.path.join(helpers.DATA, "A.fa")
    if not os.path.exists(fa):
        pytest.skip("tests/_data not staged")
    helpers.assert_summary(digests.summarize(dbg_oracle.run_oracle(fa, 301)), helpers.load_json("kat_A.json"))
