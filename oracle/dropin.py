#!/usr/bin/env python3
"""
TEST INFRASTRUCTURE ONLY -- end-to-end drop-in check (build container only: needs /root/reference).

Runs the reference's complete `Assembler.run()` (NBP graph, forward-strand selection, joining, trimming,
origin assignment: `lib/Assembler.py:365-963`, bubbles off) twice:

  reference : the unmodified `Assembler(fasta, k, ...).run(...)`
  drop-in   : the SAME reference class, but
                * `__init__` (`lib/Assembler.py:80-256`) is replaced by the attributes that
                  `superpang_b200.assembler.Assembler` fills from the C-ABI (`ksize`, `seqDict`, `ref2name`,
                  `vertex2coords`, `seqPaths`, `seqLimits`, `seqLimitsPerSeq`, `DBG` built from `spb_get_edges`);
                * the NBP-collection block of `run()` (`lib/Assembler.py:297-363`) is replaced, by a textual patch of
                  the reference source applied IN MEMORY at test time (the patch INTEGRATION.md describes), with
                  the NBPs / NBP2seq / isCycle that `collect_nbps()` returns for the same component.
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


def _patched_run(Assembler):
    """The reference's run() with the NBP-collection block replaced (returned as a plain function; it is bound to an
    instance of the ORIGINAL class so that the reference's multiprocessing pool can still pickle its classmethods)."""
    import inspect
    import textwrap
    import superpang.lib.Assembler as mod
    src = textwrap.dedent(inspect.getsource(Assembler.run))
    lines = src.split("\n")
    start = next(i for i, l in enumerate(lines) if "### Reconstruct non-branching paths" in l)
    end = next(i for i, l in enumerate(lines) if l.strip() == "NBPs = list(NBP2seq)")
    indent = lines[start][:len(lines[start]) - len(lines[start].lstrip())]
    patch = [indent + "_c = self._b200_components[ic]",
             indent + "NBPs = list(_c.NBPs); NBP2seq = dict(_c.NBP2seq); isCycle = _c.isCycle"]
    new_src = "\n".join(lines[:start] + patch + lines[end + 1:])
    ns = dict(vars(mod))
    exec(compile(new_src, "<reference run() with the INTEGRATION.md patch>", "exec"), ns)
    return ns["run"]


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
    """device_rows=True additionally serves `get_seqPaths_NBPs` (`lib/Assembler.py:1121-1139`) from the
    per-sequence NBP sets computed on the device (`spb_get_seq_nbps`) instead of the reference's substring tests."""
    Assembler = refrun.import_reference()
    from graph_tool.all import Graph
    from superpang_b200.assembler import Assembler as B200Assembler
    B = B200Assembler(fasta, k, threads, None, True, 1, engine=engine)
    import types
    A = object.__new__(Assembler)
    A.run = types.MethodType(_patched_run(Assembler), A)
    A.ksize, A.seqDict, A.ref2name, A.includedSeqs = B.ksize, B.seqDict, B.ref2name, B.includedSeqs
    A.vertex2coords = B.vertex2coords
    A.seqPaths = dict(B.seqPaths)
    A.seqLimits = B.seqLimits
    A.seqLimitsPerSeq = B.seqLimitsPerSeq
    A.DBG = Graph(directed=False)
    A.DBG.add_edge_list(B.edges)
    A.DBG.vp.comp = A.DBG.new_vertex_property("int16_t")
    A.DBG.ep.efilt = A.DBG.new_edge_property("bool")
    A._b200_components = B.collect_nbps()
    if device_rows:
        assert threads == 1                                   # the closure below is not meant to cross a fork pool
        sets = B.seq_paths_nbps()                             # {name: {vertex tuple of every raw NBP the sequence contains}}

        def get_seqPaths_NBPs(i):
            ROW1_CALLS[0] += 1
            names, seqPaths, psets, NBP2seq, vertex2NBP, seqDict = Assembler.get_multiprocessing_globals()
            mine = sets.get(names[i], ())
            return {nv for pair in psets.values() for nv in pair if vertex2NBP[nv] in mine}
        A.get_seqPaths_NBPs = get_seqPaths_NBPs
    with contextlib.redirect_stdout(io.StringIO()):
        return A.run(max_threads=threads, debug=True, **RUN_ARGS)


if __name__ == "__main__":
    import ctypes
    import json
    from superpang_b200 import _capi
    fasta, k = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 301
    lib = _capi._declare(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libspb_emu.so")))
    ref = run_reference_full(fasta, k, threads=8)
    got = run_dropin_full(fasta, k, _capi.Engine(lib=lib), threads=8)
    print(json.dumps(dict(reference=contigs_digest(ref), dropin=contigs_digest(got), identical=canonical(ref) == canonical(got)), indent=1))
