"""GPU tier, BENCHMARK scale: BASELINE config 3 (100 genomes x 5 Mbp, k = 301) -- the configuration `bench.py` times.

The oracle cannot run 5e8 instances, so parity at this size is pinned by:
  * two different dedup algorithms -- the default path the benchmark takes (partitioned, shared-memory tables in `snp`
    mode; anchor + extend in `block` mode) and the single global table (`SPB_DEDUP=global`) -- must give byte-identical
    result buffers (checksummed on the device: vertex table, edges, per-instance ids, NBP vertex lists / sequences /
    origins);
  * N_v equals an independent sort-count that shares no code with the library (`digest.independent_vertex_count`);
  * every vertex is an NBP end or interior to exactly one NBP pair;
  * sampled NBP windows re-k-merise to the vertices the NBP lists (canonical k-mer of the vertex' first appearance).
Both algorithms are checked against the oracle / the reference goldens at small sizes in test_gpu_parity.py.
"""
import numpy as np
import pytest

from superpang_b200 import _capi, digest, synth
from superpang_b200.assembler import concat_sequences

pytestmark = pytest.mark.gpu

K = 301
ENV_KEYS = ("SPB_DEDUP", "SPB_ANCHOR", "SPB_BUCKET_BITS", "SPB_CHUNK_BITS", "SPB_SMEM_AVG", "SPB_PARTITION")


def _run(monkeypatch, env, dbuf, off, graph=True, k=None):
    for k_ in ENV_KEYS:
        monkeypatch.delenv(k_, raising=False)
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    eng = _capi.Engine(device=0)
    eng.load(dbuf, off, k or K)
    eng.build_dbg()
    eng.compact()
    stats = eng.stage_stats()
    return eng, digest.result_digest(eng, graph=graph), stats


@pytest.mark.parametrize("mode", ["snp", "block"])
def test_config3_two_algorithms_and_invariants(monkeypatch, mode):
    import torch
    names, seqs = synth.genome_set(100, 5_000_000, 0.96, seed=2, mode=mode)
    buf, off = concat_sequences(seqs)
    n_inst = synth.n_instances(seqs, K)
    del seqs
    dev = torch.device("cuda", 0)
    dbuf = torch.from_numpy(buf).to(dev)

    # ---- (1) the benchmark's default path vs the global table: identical buffers
    eng, d_default, st_default = _run(monkeypatch, {}, dbuf, off)
    assert d_default["counts"]["n_inst"] == n_inst
    if mode == "snp":        # the path the benchmark line is measured on must be the one checked here
        assert "smem_tables" in st_default or "partition2" in st_default, st_default
    eng.close()
    torch.cuda.empty_cache()
    eng_g, d_global, st_global = _run(monkeypatch, {"SPB_DEDUP": "global"}, dbuf, off)
    assert "k_extract_insert_ms" in st_global and "switched_to_links" not in st_global, st_global
    assert d_default == d_global, {k_: (d_default[k_], d_global[k_]) for k_ in d_default if d_default[k_] != d_global[k_]}
    eng_g.close()
    torch.cuda.empty_cache()

    # ---- (2) N_v from an independent sort-count
    nv = digest.independent_vertex_count(dbuf, off, K)
    torch.cuda.empty_cache()
    assert nv == d_default["counts"]["n_vertices"]

    # ---- (3) structure of the NBP set (default path again; buffers stay on the device)
    eng, d_again, _ = _run(monkeypatch, {}, dbuf, off, graph=False)
    assert all(d_again[k_] == d_default[k_] for k_ in d_again), "the default path is not deterministic"
    c = eng.counts
    L, h = eng.lib, eng.h
    voff = digest._filled(eng, c.n_nbp + 1, torch.int64, lambda p: L.spb_get_nbps(h, p, None, None, None))
    verts = digest._filled(eng, c.nbp_vertices, torch.int32, lambda p: L.spb_get_nbps(h, None, p, None, None))
    soff = digest._filled(eng, c.n_nbp + 1, torch.int64, lambda p: L.spb_get_nbp_seqs(h, p, None))
    assert bool(((soff[1:] - soff[:-1]) == (voff[1:] - voff[:-1]) + (K - 1)).all())
    interior = torch.ones(verts.numel(), dtype=torch.bool, device=dev)
    interior[voff[:-1]] = False
    interior[voff[1:] - 1] = False
    cnt = torch.bincount(verts[interior].long(), minlength=int(c.n_vertices))
    is_end = torch.zeros(int(c.n_vertices), dtype=torch.bool, device=dev)
    is_end[verts[~interior].long()] = True
    assert bool(((cnt == 2) | ((cnt == 0) & is_end)).all()), "a vertex is neither an NBP end nor interior to exactly one NBP pair"
    assert int((cnt == 2).sum().item()) + int(is_end.sum().item()) == int(c.n_vertices)
    del cnt, interior, is_end
    # origins: non-empty, in range
    ooff = digest._filled(eng, c.n_nbp + 1, torch.int64, lambda p: L.spb_get_nbp_origins(h, p, None))
    oidx = digest._filled(eng, c.n_origin_pairs, torch.int32, lambda p: L.spb_get_nbp_origins(h, None, p))
    assert bool((ooff[1:] > ooff[:-1]).all()) and int(oidx.max().item()) < len(names) and int(oidx.min().item()) >= 0

    # ---- (4) sampled NBP windows re-k-merise to their vertices
    seqs_dev = digest._filled(eng, c.nbp_bases, torch.uint8, lambda p: L.spb_get_nbp_seqs(h, None, p))
    v2c = digest._filled(eng, 2 * c.n_vertices, torch.int32, lambda p: L.spb_get_vertex2coords(h, p)).view(-1, 2)
    rng = np.random.default_rng(0)
    comp_tab = bytes.maketrans(b"ACGT", b"TGCA")
    voff_h, soff_h = voff.cpu().numpy(), soff.cpu().numpy()
    nlen = voff_h[1:] - voff_h[:-1]
    picks = rng.choice(len(nlen), size=min(400, len(nlen)), replace=False)
    checked = 0
    for i in picks:
        for r in rng.integers(0, nlen[i], size=5):
            v = int(verts[int(voff_h[i]) + int(r)].item()) & 0xFFFFFFFF
            ref_i, idx = (int(x) & 0xFFFFFFFF for x in v2c[v].tolist())
            p = int(off[ref_i]) + idx
            km = bytes(seqs_dev[int(soff_h[i]) + int(r):int(soff_h[i]) + int(r) + K].cpu().numpy())
            src = buf[p:p + K].tobytes()
            assert km == src or km == src.translate(comp_tab)[::-1], (i, r, v)
            checked += 1
    assert checked >= 5
    eng.close()


def test_config5_k_sweep_two_algorithms(monkeypatch):
    """BASELINE config 5 (k = 101 / 501 on 100 x 5 Mbp): the default path and the single global table -- two different
    algorithms -- must give byte-identical device digests at full scale for the other k values of the sweep too, and N_v
    must equal the independent sort-count."""
    import torch
    names, seqs = synth.genome_set(100, 5_000_000, 0.96, seed=2, mode="snp")
    buf, off = concat_sequences(seqs)
    del seqs
    dev = torch.device("cuda", 0)
    dbuf = torch.from_numpy(buf).to(dev)
    for k in (101, 501):
        eng, d_default, st = _run(monkeypatch, {}, dbuf, off, k=k)
        eng.close()
        torch.cuda.empty_cache()
        eng_g, d_global, st_g = _run(monkeypatch, {"SPB_DEDUP": "global"}, dbuf, off, k=k)
        eng_g.close()
        torch.cuda.empty_cache()
        assert "k_extract_insert_ms" in st_g, st_g
        assert d_default == d_global, (k, {k_: (d_default[k_], d_global[k_]) for k_ in d_default if d_default[k_] != d_global[k_]})
        assert digest.independent_vertex_count(dbuf, off, k) == d_default["counts"]["n_vertices"], k
