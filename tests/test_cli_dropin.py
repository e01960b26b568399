"""BASELINE config 1 / SURVEY 8(b): the reference CLI (`scripts/SuperPang.py`) with the B200 `Assembler` bound at its only
call site (`:199`) must write byte-identical output files -- `NBPs.fasta`, `NBP2origins.tsv`, `graph.fastg`,
`assembly.info.tsv` (+ `assembly.fasta`, `graph.NBP2origins.csv`) -- and accept `-k`, `-t`, `--lowmem` unchanged.

CPU tier: small slices of the bundled genomes, reference CLI vs drop-in CLI on the kernel-logic emulator (needs the staged
reference, oracle/_ref/superpang, i.e. a build container or a box that received the snapshot).
GPU tier: the reference's own smoke test (`scripts/test_SuperPang.py:23-28`, "fast" pair, `-i 1`) through the CUDA library,
against the sha256 of the files the unmodified reference wrote (tests/golden/cli_fast.json, oracle/make_golden_cli.py).
graph-tool / mOTUlizer / minimap2 are absent from this image: both sides run under the same import shims with `-i 1 -b 1`."""
import json
import os

import pytest

import helpers
import make_golden_cli as cli
import refrun
from superpang_b200.assembler import read_fasta

GENOMES = os.path.join(helpers.DATA, "genomes")


def _need_staged():
    if not refrun.available():
        pytest.skip("reference not staged (oracle/_ref/superpang)")
    if not all(os.path.exists(os.path.join(GENOMES, f)) for f in cli.FAST):
        pytest.skip("tests/_data/genomes not staged")


def _slices(tmp_path, n_rec, length):
    out = []
    for f in cli.FAST:
        d = read_fasta(os.path.join(GENOMES, f))
        p = tmp_path / f
        with open(p, "w") as o:
            for i, (n, s) in enumerate(d.items()):
                if i < n_rec:
                    o.write(f">{n}\n{s[:length]}\n")
        out.append(str(p))
    return out


@pytest.mark.parametrize("k,threads,extra", [(301, 1, ()), (301, 2, ("--lowmem",)), (101, 2, ())], ids=["k301-t1", "k301-t2-lowmem", "k101-t2"])
def test_cli_dropin_matches_reference_cpu(tmp_path, emu_lib, k, threads, extra):
    _need_staged()
    genomes = _slices(tmp_path, 4, 25000)
    ref, got = str(tmp_path / "ref"), str(tmp_path / "dropin")
    cli.run_cli("reference", genomes, ref, threads=threads, extra=("-k", str(k)))
    cli.run_cli("dropin", genomes, got, threads=threads, extra=("-k", str(k)) + tuple(extra), emu=True)
    for f in cli.FILES:
        assert open(os.path.join(ref, f), "rb").read() == open(os.path.join(got, f), "rb").read(), f
    assert os.path.getsize(os.path.join(got, "NBPs.fasta")) > 0


@pytest.mark.gpu
def test_cli_dropin_fast_pair_gpu(tmp_path):
    """The reference's own smoke test through the CUDA library; -t and --lowmem passed through the real CLI."""
    _need_staged()
    want = json.load(open(os.path.join(helpers.GOLD, "cli_fast.json")))
    out = str(tmp_path / "out")
    cli.run_cli("dropin", [os.path.join(GENOMES, f) for f in want["genomes"]], out, threads=8, extra=("--lowmem",))
    got = cli.digests(out)
    for f in cli.FILES:
        assert got[f] == want["sha256"][f], f
