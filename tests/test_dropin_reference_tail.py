"""CPU tier, build container only: the drop-in claim end to end.  The reference's complete `Assembler.run()`
(host tail `lib/Assembler.py:365-963`, bubbles off) is executed on top of the B200 path's DBG + NBPs (kernels run by
the host emulator here) through the in-memory patch of `oracle/dropin.py`, and must produce exactly the final
contigs of the unmodified reference.  Skipped wherever /root/reference is absent."""
import os

import pytest

import cases
import refrun
from superpang_b200 import _capi

pytestmark = pytest.mark.skipif(not refrun.available(), reason="reference tree not present")


@pytest.mark.parametrize("case", cases.crafted_cases(k=31), ids=lambda c: c[0])
def test_reference_tail_on_b200_head(emu_lib, case, tmp_path):
    import dropin
    name, k, recs = case
    fa = os.path.join(tmp_path, "in.fa")
    cases.write_fasta(fa, recs)
    try:
        ref = dropin.run_reference_full(fa, k)
    except AssertionError:                          # the reference asserts on this input: the drop-in must refuse
        with pytest.raises(_capi.DegenerateInputError):
            dropin.run_dropin_full(fa, k, _capi.Engine(lib=emu_lib))
        return
    got = dropin.run_dropin_full(fa, k, _capi.Engine(lib=emu_lib))
    assert dropin.canonical(ref) == dropin.canonical(got)
    assert dropin.contigs_digest(ref) == dropin.contigs_digest(got)
    # SURVEY 8(f) row 1 wired in as well: get_seqPaths_NBPs answered from the device-side per-sequence NBP sets
    from superpang_b200 import dropin as product
    product._ROW1["joined"] = 0
    got = dropin.run_dropin_full(fa, k, _capi.Engine(lib=emu_lib), device_rows=True)
    assert dropin.canonical(ref) == dropin.canonical(got)
    if name == "inverted_repeat":                   # the one crafted case whose tail joins NBPs (bubble popping is off here):
        assert product._ROW1["joined"] >= 1         # their sequences came from spb_path_sequences


def test_reference_tail_on_b200_head_synthetic(emu_lib, tmp_path):
    import dropin
    from superpang_b200 import synth
    names, seqs = synth.genome_set(5, 30000, 0.96, seed=77, mode="block")
    fa = os.path.join(tmp_path, "syn.fa")
    synth.write_fasta(fa, names, seqs)
    ref = dropin.run_reference_full(fa, 301)
    got = dropin.run_dropin_full(fa, 301, _capi.Engine(lib=emu_lib))
    assert dropin.canonical(ref) == dropin.canonical(got)
    before = dropin.ROW1_CALLS[0]
    got = dropin.run_dropin_full(fa, 301, _capi.Engine(lib=emu_lib), device_rows=True)
    assert dropin.canonical(ref) == dropin.canonical(got)
    assert dropin.ROW1_CALLS[0] >= before + 5           # the reference's run() did go through the device-backed function
