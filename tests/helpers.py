"""Shared helpers for the parity tests."""
import gzip
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "tests", "_data")        # staged reference test data (git-ignored, travels to the GPU box)


def load_crafted():
    with gzip.open(os.path.join(GOLD, "crafted.json.gz"), "rt") as f:
        return json.load(f)["cases"]


def golden_to_result(case):
    r = case["result"]
    return dict(
        k=case["k"], names=r["names"], n_inst=r["n_inst"], ncomp=r["ncomp"],
        vertex2coords=np.asarray(r["vertex2coords"], dtype=np.uint32).reshape(-1, 2),
        edges=np.asarray(r["edges"], dtype=np.uint32).reshape(-1, 2),
        seqPaths={n: np.asarray(v, dtype=np.uint32).reshape(-1, 2) for n, v in r["seqPaths"].items()},
        seqLimits=list(r["seqLimits"]),
        nbps=[(tuple(v), s, tuple(o), c, cy) for v, s, o, c, cy in r["nbps"]],
    )


def load_json(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def real_slice_records():
    with gzip.open(os.path.join(GOLD, "real_slice.fa.gz"), "rt") as f:
        txt = f.read().strip().lstrip(">")
    return dict((rec.split("\n", 1)[0], rec.split("\n", 1)[1].replace("\n", "")) for rec in txt.split("\n>"))


def synth_seqdict(n_genomes, length, ani, seed, mode):
    from superpang_b200 import synth
    names, seqs = synth.genome_set(n_genomes, length, ani, seed=seed, mode=mode)
    return {n: s.tobytes().decode() for n, s in zip(names, seqs)}


SUMMARY_KEYS = ("n_inst", "n_vertices", "n_edges", "n_comp", "n_nbp", "nbp_bases", "nbp_len_min", "nbp_len_max",
                "n_nbp_2v", "n_cycle_nbp", "sha_vertex2coords", "sha_edges", "sha_seqlimits", "sha_seqpaths",
                "D_nbpv", "D_seq", "D_ori")


def assert_summary(got, want):
    for key in SUMMARY_KEYS:
        assert got[key] == want[key], f"{key}: {got[key]} != {want[key]}"
