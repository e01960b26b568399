"""Host-side pieces of the measurement (no GPU): the sliced result digest is additive over parts, the byte models of bench.py
are what DESIGN.md states, and the device-side generator of the config-4 workload is deterministic."""
import os
import sys

import numpy as np
import torch

import helpers
from superpang_b200 import digest, synth

sys.path.insert(0, helpers.ROOT)


def test_checksum_is_additive_over_parts():
    rng = np.random.default_rng(7)
    for dt in (torch.uint8, torch.int32, torch.int64):
        x = torch.from_numpy(rng.integers(0, 200, size=100_003)).to(dt)
        whole = digest.checksum(x)
        cuts = [0, 1, 4096, 65_537, 100_003]
        parts = sum(digest.checksum(x[a:b], base=a) for a, b in zip(cuts[:-1], cuts[1:])) & 0xFFFFFFFFFFFFFFFF
        assert whole == parts
        assert digest.checksum(torch.roll(x, 1)) != whole          # order-sensitive


def test_byte_models():
    import bench
    c = dict(n_inst=10, n_bases=16, n_vertices=8, n_edges=9, nbp_bases=100, nbp_vertices=20)
    assert bench.algorithmic_bytes(10, 8, 100, k=301) == 10 * (16 * 10 + 24) + 8 * (8 * 10 + 32) + 100      # SURVEY 8(d), W = 10 at k = 301
    assert bench.KERNEL_BYTES["k_roll_partition_ms"](c) == 8 * 10 + 8                                         # 8 B record out + 2 x 2 bits per base in
    assert bench.KERNEL_BYTES["k_partition2_ms"](c) == 160
    assert set(bench.WORKLOADS) >= {"c2", "c3", "c4s"}
    assert bench.WORKLOADS["c4s"]["n_genomes"] * bench.WORKLOADS["c4s"]["length"] < (1 << 32) - 4096          # the position limit of this build


def test_device_generator_is_deterministic_and_diverged():
    a, off = synth.genome_set_device(3, 20_000, 0.96, 4, torch.device("cpu"))
    b, _ = synth.genome_set_device(3, 20_000, 0.96, 4, torch.device("cpu"))
    assert torch.equal(a, b) and list(off) == [0, 20_000, 40_000, 60_000]
    assert set(np.unique(a.numpy()).tolist()) <= set(b"ACGT")
    g = a.numpy().reshape(3, 20_000)
    ident = (g[0] == g[1]).mean()
    assert 0.94 < ident < 0.98                                     # two genomes at p = 0.02 each: ~0.96 pairwise identity
