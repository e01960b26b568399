"""NBP graph: successors by k-overlap and reverse partners, against the definition applied to the
reference's NBP sequences (`lib/Assembler.py:365-384`).  CPU tier through the emulator; GPU tier through the library."""
import numpy as np
import pytest

import helpers
from superpang_b200 import _capi
from superpang_b200.assembler import Assembler


def check(A, k):
    raw = A.raw_nbps()
    off, succ, rev = A._eng.nbp_graph()
    seqs = [r[1] for r in raw]
    verts = [r[0] for r in raw]
    start = {}
    for j, s in enumerate(seqs):
        start.setdefault(s[:k], []).append(j)
    for i, s in enumerate(seqs):
        want = sorted(start.get(s[-k:], []))
        got = sorted(succ[int(off[i]):int(off[i + 1])].tolist())
        assert got == want, (i, got, want)
        if not raw[i][4]:                                   # reverse partner: the reversed vertex tuple
            assert verts[int(rev[i])] == verts[i][::-1]
        else:                                               # broken cycle: partner runs the ring the other way round
            assert set(verts[int(rev[i])]) == set(verts[i]) and int(rev[int(rev[i])]) == i


@pytest.mark.parametrize("case", [c for c in helpers.load_crafted() if c["k"] == 31 and not c.get("degenerate")], ids=lambda c: c["name"])
def test_nbp_graph_emu(emu_lib, case):
    check(Assembler(dict(map(tuple, case["records"])), case["k"], engine=_capi.Engine(lib=emu_lib)), case["k"])


def test_nbp_graph_emu_synthetic(emu_lib):
    sd = helpers.synth_seqdict(6, 30000, 0.96, 12, "block")
    check(Assembler(sd, 31, engine=_capi.Engine(lib=emu_lib)), 31)


@pytest.mark.gpu
def test_nbp_graph_gpu():
    sd = helpers.synth_seqdict(6, 30000, 0.96, 12, "block")
    check(Assembler(sd, 31, device=0), 31)
    check(Assembler(helpers.real_slice_records(), 301, device=0), 301)
