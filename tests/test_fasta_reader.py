"""SURVEY 8(f) row 3 (reader half): the library's native FASTA reader against the Python mirror of the reference's
read_fasta (superpang_b200/assembler.py) and, where the reference tree is present, against the reference function itself."""
import os

import numpy as np
import pytest

import refrun
from superpang_b200 import _capi
from superpang_b200.assembler import read_fasta

TEXTS = {
    "plain": ">a desc\nACGT\nACGT\n>b\nTTTT\n",
    "case_u_ambig": ">a\nacgu\nRYKM\nswbd\nhvN\n>b\nAC.G-T  \n",
    "gaps1": ">a\nACGTNACGTNNACGT\n>b\nNNACGTNN\n>c\nNNNN\n",
    "leading_blank": "\n\n>>a\nACGT\n\n>b x y\nAC GT\n\n\n",
    "tabs_cr": ">a\tb c\nAC\tGT\r\nAC\n>d\nGG\n",
    "one_record": ">only\nACGTACGT",
    "nsplit_named": ">a_Nsplit_0\nACGT\n>b\nACNNGT\n",
    # CRLF / lone-CR files: the reference opens in text mode (universal newlines), so names must not keep a '\r' 
    "crlf": ">s1\r\nACGTNACGT\r\nAC\r\n>s2 desc\r\nGGNNTT\r\n",
    "lone_cr": ">s1\rACGT\rAC\r>s2\rGG\r",
}
BAD = {
    "illegal": ">a\nACGTX\n",
    "duplicate": ">a\nACGT\n>a\nGGGG\n",
    "no_sequence_line": ">a",
}


def _write(tmp_path, name, txt):
    p = os.path.join(tmp_path, name + ".fa")
    with open(p, "w", newline="") as f:
        f.write(txt)
    return p


def _native(emu_lib, path, gap):
    names, buf, off = _capi.read_fasta_native(path, gap, lib=emu_lib)
    b = buf.tobytes().decode()
    return {n: b[int(off[i]):int(off[i + 1])] for i, n in enumerate(names)}, names


@pytest.mark.parametrize("gap", [1, 2, 3])
@pytest.mark.parametrize("name", sorted(TEXTS))
def test_native_reader_matches_python_mirror(emu_lib, tmp_path, name, gap):
    p = _write(tmp_path, name, TEXTS[name])
    want = read_fasta(p, gap_size=gap)
    got, order = _native(emu_lib, p, gap)
    assert got == want and order == list(want)


@pytest.mark.parametrize("name", sorted(BAD))
def test_native_reader_refuses_what_the_reference_refuses(emu_lib, tmp_path, name):
    p = _write(tmp_path, name, BAD[name])
    with pytest.raises(Exception):
        read_fasta(p)
    with pytest.raises(ValueError):
        _capi.read_fasta_native(p, 1, lib=emu_lib)


def test_native_reader_random_texts(emu_lib, tmp_path):
    rng = np.random.default_rng(5)
    alphabet = list("ACGTacgtNnUuRYKMSWBDHV.- \n")
    for it in range(60):
        recs = []
        for r in range(int(rng.integers(1, 5))):
            body = "".join(rng.choice(alphabet, size=int(rng.integers(1, 80))))
            recs.append(f">r{it}_{r} d\n{body}")
        p = _write(tmp_path, f"rnd{it}", "\n".join(recs) + "\n")
        for gap in (1, 2):
            try:
                want = read_fasta(p, gap_size=gap)
            except Exception:
                with pytest.raises(ValueError):
                    _capi.read_fasta_native(p, gap, lib=emu_lib)
                continue
            got, order = _native(emu_lib, p, gap)
            assert got == want and order == list(want)


@pytest.mark.skipif(not refrun.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", sorted(TEXTS))
def test_native_reader_matches_reference_function(emu_lib, tmp_path, name):
    refrun.import_reference()
    from superpang.lib.utils import read_fasta as ref_read
    p = _write(tmp_path, name, TEXTS[name])
    for gap in (1, 2):
        want = ref_read(p, ambigs="as_Ns", Ns="split", gap_size=gap)
        got, order = _native(emu_lib, p, gap)
        assert got == want and order == list(want)
