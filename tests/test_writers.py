"""SURVEY 8(f) row 3, writer half: the native FASTA / origin-table writers produce byte-identical files to the reference's
own code (`lib/utils.py:62-65`, `scripts/SuperPang.py:284-294`); sequences may come straight from device buffers."""
import os
import sys

import numpy as np
import pytest

import helpers
from superpang_b200 import _capi, writers
from superpang_b200.assembler import Assembler

REF = "/root/reference/src"


def _ref_write_fasta(seqDict, path):
    """The reference's write_fasta: imported live where the reference tree exists (build container), else its two lines."""
    if os.path.isdir(os.path.join(REF, "superpang")):
        sys.path.insert(0, os.path.join(helpers.ROOT, "oracle", "shims"))
        sys.path.insert(0, REF)
        try:
            from superpang.lib.utils import write_fasta
            return write_fasta(seqDict, path)
        except Exception:
            pass
    with open(path, "w") as f:                       # lib/utils.py:62-65
        for name, seq in seqDict.items():
            f.write(f">{name}\n{seq}\n")


def _ref_origin_table(path, first, node_names, node_origins, name2bin, n_genomes):
    with open(path, "w") as f:                       # scripts/SuperPang.py:284-294
        f.write(f"{first}\tSourceContigs\tSourceGenomes\tNumberOfSourceGenomes\n")
        for node, origins in zip(node_names, node_origins):
            origs = ",".join(origins)
            bins = [name2bin[n] for n in origins]
            lbins = len(set(bins))
            f.write(f"{node}\t{origs}\t{','.join(bins)}\t{lbins}/{n_genomes}\n")


def test_write_fasta_matches_reference(emu_lib, tmp_path):
    rng = np.random.default_rng(3)
    d = {f"NODE_Sc{i}-0-noinfo_length_{n}_cov_1.5_tag_noinfo;": "".join("ACGT"[x] for x in rng.integers(0, 4, n))
         for i, n in enumerate([1, 301, 5000, 77, 0, 12345])}
    a, b = os.path.join(tmp_path, "a.fa"), os.path.join(tmp_path, "b.fa")
    writers.write_fasta(d, a, lib=emu_lib)
    _ref_write_fasta(d, b)
    assert open(a, "rb").read() == open(b, "rb").read()
    writers.write_fasta({}, a, lib=emu_lib)
    assert open(a, "rb").read() == b""


def test_cuda_library_writes_without_a_device(tmp_path):
    """The writers are host code inside the product library too (the CLI drop-in binds them): they must work on host
    buffers whether or not a GPU is present."""
    d = {"a;": "ACGT" * 100, "b": "", "c": "T"}
    a, b = os.path.join(tmp_path, "a.fa"), os.path.join(tmp_path, "b.fa")
    writers.write_fasta(d, a)                       # lib=None: libsuperpang_b200.so
    _ref_write_fasta(d, b)
    assert open(a, "rb").read() == open(b, "rb").read()


def test_origin_table_matches_reference(emu_lib, tmp_path):
    name2bin = {f"g{g}|c{c}": f"g{g}" for g in range(5) for c in range(3)}
    names = list(name2bin)
    rng = np.random.default_rng(5)
    node_names = [f"NODE_Sc{i}-0-core_length_{10 * i}_cov_2.0_tag_core;" for i in range(40)]
    node_origins = [tuple(sorted(rng.choice(names, size=int(rng.integers(0, 9)), replace=False))) for _ in node_names]
    a, b = os.path.join(tmp_path, "a.tsv"), os.path.join(tmp_path, "b.tsv")
    writers.write_origins_table(a, "Node", node_names, node_origins, name2bin, 5, lib=emu_lib)
    _ref_origin_table(b, "Node", node_names, node_origins, name2bin, 5)
    assert open(a, "rb").read() == open(b, "rb").read()


def _check_engine_fasta(engine, tmp_path):
    from superpang_b200 import synth
    names, seqs = synth.genome_set(4, 15000, 0.96, seed=33, mode="block")
    sd = {n: s.tobytes().decode() for n, s in zip(names, seqs)}
    A = Assembler(sd, 31, engine=engine)
    A.compact()
    path = os.path.join(tmp_path, "nbps.fa")
    n = A._eng.write_nbp_fasta(path)
    soff, seq = A._eng.nbp_seqs()
    b = seq.tobytes()
    want = "".join(f">NBP_{i}\n{b[int(soff[i]):int(soff[i + 1])].decode()}\n" for i in range(n)).encode()
    assert n == A.counts["n_nbp"] and n > 10
    assert open(path, "rb").read() == want


def test_engine_writes_nbp_fasta_emulator(emu_lib, tmp_path):
    _check_engine_fasta(_capi.Engine(lib=emu_lib), tmp_path)


@pytest.mark.gpu
def test_engine_writes_nbp_fasta_from_device_buffers(tmp_path):
    _check_engine_fasta(_capi.Engine(device=0), tmp_path)
