"""CPU tier, build container only: the oracle restatement against the unmodified reference run live
(oracle/refrun.py).  Skipped wherever /root/reference is absent (e.g. on the GPU box)."""
import os

import pytest

import cases
import dbg_oracle
import digests
import refrun

pytestmark = pytest.mark.skipif(not refrun.available(), reason="reference tree not present")


@pytest.mark.parametrize("case", cases.crafted_cases(k=35, seed=17), ids=lambda c: c[0])
def test_live_reference_matches_oracle(case, tmp_path):
    name, k, recs = case
    fa = os.path.join(tmp_path, "in.fa")
    cases.write_fasta(fa, recs)
    try:
        ref = refrun.run_reference(fa, k)
    except AssertionError:                          # orientation-ambiguous input: the oracle must assert as well
        with pytest.raises(AssertionError):
            dbg_oracle.run_oracle(fa, k)
        return
    digests.assert_same(ref, dbg_oracle.run_oracle(fa, k))
