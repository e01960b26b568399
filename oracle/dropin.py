#!/usr/bin/env python3
"""
TEST INFRASTRUCTURE ONLY -- end-to-end drop-in check (build container only: needs /root/reference).

Runs the reference's complete `Assembler.run()` (NBP graph, forward-strand selection, joining, trimming,
origin assignment: `lib/Assembler.py:365-963`, bubbles off) twice:

  reference : the unmodified `Assembler(fasta, k, ...).run(...)`
  drop-in   : the product's `superpang_b200.assembler.Assembler(...).run(...)` (superpang_b200/dropin.py): DBG + NBP
              collection from the C-ABI, host tail = the installed reference's own `run()` with the NBP-collection block
              (`lib/Assembler.py:297-363`) replaced in memory
and compares the final contigs (seq, tseq, cov, origins, successors structure).  Nothing of the reference is
copied into the repository; the product never imports this file.
"""
import hashlib
import io
import contextlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import refrun  # noqa: E402

ROW1_CALLS = [0]            # how often the device-backed get_seqPaths_NBPs was called (tests check that it is used)
RUN_ARGS = dict(minlen=0, mincov=0, bubble_identity_threshold=1, genome_assignment_threshold=1)


def contigs_digest(contigs):
    tseq = hashlib.sha256("\n".join(sorted(c.tseq for c in contigs.values() if c.tseq)).encode()).hexdigest()
    seq = hashlib.sha256("\n".join(sorted(c.seq for c in contigs.values())).encode()).hexdigest()
    ori = hashlib.sha256("\n".join(sorted(c.seq + "\t" + ",".join(sorted(c.origins)) for c in contigs.values())).encode()).hexdigest()
    return dict(n_contigs=len(contigs), n_tseq=sum(1 for c in contigs.values() if c.tseq),
                tseq_bases=sum(len(c.tseq) for c in contigs.values()), seq_bases=sum(len(c.seq) for c in contigs.values()),
                D_final_tseq=tseq, D_final_seq=seq, D_final_ori=ori)


def canonical(contigs):
    """Order-independent view incl. the successor structure (successors named by their sequences)."""
    return sorted((c.seq, c.tseq, round(float(c.cov), 6), tuple(sorted(c.origins)),
                   tuple(sorted(contigs[s].seq for s in c.successors))) for c in contigs.values())


def run_reference_full(fasta, k, threads=1):
    Assembler = refrun.import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        return Assembler(fasta, k, threads, None, True, 1).run(max_threads=threads, debug=True, **RUN_ARGS)


def run_dropin_full(fasta, k, engine, threads=1, device_rows=False):
    """The PRODUCT's drop-in (`superpang_b200.assembler.Assembler(...).run(...)`, see superpang_b200/dropin.py) on `engine`.
    device_rows=True additionally serves `get_seqPaths_NBPs` (`lib/Assembler.py:1121-1139`) from the
    per-sequence NBP sets computed on the device (`spb_get_seq_nbps`) instead of the reference's substring tests."""
    refrun.import_reference()                                  # puts the staged reference + shims on sys.path
    from superpang_b200 import dropin as product
    from superpang_b200.assembler import Assembler as B200Assembler
    B = B200Assembler(fasta, k, threads, None, True, 1, engine=engine)
    before = product._ROW1["calls"]
    with contextlib.redirect_stdout(io.StringIO()):
        out = product.run(B, max_threads=threads, debug=True, device_rows=device_rows, **RUN_ARGS)
    ROW1_CALLS[0] += product._ROW1["calls"] - before
    return out


if __name__ == "__main__":
    import ctypes
    import json
    from superpang_b200 import _capi
    fasta, k = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 301
    lib = _capi._declare(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libspb_emu.so")))
    ref = run_reference_full(fasta, k, threads=8)
    got = run_dropin_full(fasta, k, _capi.Engine(lib=lib), threads=8)
    print(json.dumps(dict(reference=contigs_digest(ref), dropin=contigs_digest(got), identical=canonical(ref) == canonical(got)), indent=1))
