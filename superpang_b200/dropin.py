"""
Drop-in seam: `Assembler.run()` for the B200 path (`scripts/SuperPang.py:199` is the only call site of the reference).

The DBG build and the NBP collection (`lib/Assembler.py:80-256`, `:269-363`) come from the CUDA library through the
C-ABI.  Everything after that -- NBP graph, forward-strand selection, bubble popping, joining, trimming, origin
assignment (`lib/Assembler.py:365-963`) -- is host control flow that SURVEY 8 leaves on the host: it is NOT
re-implemented here but taken from the INSTALLED reference package at run time (`import superpang.lib.Assembler`):

  * an instance of the reference class is created without running its `__init__` and given the attributes the tail
    reads (`ksize`, `seqDict`, `ref2name`, `vertex2coords`, `seqPaths`, `seqLimits`, `seqLimitsPerSeq`, `DBG`);
  * the reference's own `run()` is used with ONE block replaced -- the NBP collection of each component
    (`lib/Assembler.py:297-363`, between the markers "### Reconstruct non-branching paths" and
    "NBPs = list(NBP2seq)") takes NBPs / NBP2seq / isCycle from the device results -- which is the patch
    INTEGRATION.md asks a maintainer to make in the source; here it is applied in memory;
  * `get_seqPaths_NBPs` (`lib/Assembler.py:1121-1139`, SURVEY 8f row 1: the quadratic sequence x NBP substring
    test of the forward-strand selection) is answered from the per-sequence NBP sets computed on the device
    (`spb_get_seq_nbps`), and the NBP graph of every component (`:365-384`: successors by equality of the last / first k
    bases) is built from the successor lists of `spb_get_nbp_graph` instead of 2 x #NBPs k-base string keys, and the
    sequences of the joined paths (`:826`, row 2) come from `spb_path_sequences`.

`main()` runs the reference's CLI (`scripts/SuperPang.py`) with this class bound at its call site and the native FASTA
writer (`spb_write_fasta`) bound in place of `lib/utils.py:write_fasta`, so `-k`, `-t`,
`--lowmem` and the output files (`NBPs.fasta`, `NBP2origins.tsv`, `graph.fastg`, `assembly.info.tsv`) are the CLI's own.
"""
import inspect
import sys
import textwrap
import types

# state read by `_seq_paths_nbps_from_device` (module level so that the reference's fork pool can pickle the function
# by name; forked workers inherit the state)
_ROW1 = {"sets": None, "ref_class": None, "calls": 0}


def reference_class():
    try:
        from superpang.lib.Assembler import Assembler as Ref
    except ModuleNotFoundError as e:
        raise RuntimeError("Assembler.run() needs the SuperPang package importable (its host tail, lib/Assembler.py:365-963, "
                           "is used as installed); only the DBG build + NBP compaction run on the GPU") from e
    return Ref


def patched_run(Ref, nbp_graph=False):
    """The reference's `run()` with the per-component NBP-collection block replaced (see the module docstring).
    nbp_graph=True also replaces the construction of the NBP graph (`lib/Assembler.py:365-384`: every NBP's first / last k
    bases as dict keys, successors by string equality) by the successor lists computed on the device (`spb_get_nbp_graph`,
    SURVEY 8f row 1).  The groups are rebuilt as Python sets with the SAME members inserted in the SAME order as the
    reference's `start2NBPs`, so that the edge order of the graph -- and with it everything downstream -- is unchanged."""
    import superpang.lib.Assembler as mod
    lines = textwrap.dedent(inspect.getsource(Ref.run)).split("\n")
    start = next(i for i, l in enumerate(lines) if "### Reconstruct non-branching paths" in l)
    end = next(i for i, l in enumerate(lines) if l.strip() == "NBPs = list(NBP2seq)")
    indent = lines[start][:len(lines[start]) - len(lines[start].lstrip())]
    patch = [indent + "_c = self._b200_components[ic]",
             indent + "NBPs = list(_c.NBPs); NBP2seq = dict(_c.NBP2seq); isCycle = _c.isCycle"]
    tail = lines[end + 1:]
    if nbp_graph:
        g0 = next(i for i, l in enumerate(tail) if "# Collect prefixes-suffixes for locating potential overlaps" in l)
        g1 = next(i for i, l in enumerate(tail) if l.strip() == "# Build graph of non-branching paths")
        gi = tail[g0][:len(tail[g0]) - len(tail[g0].lstrip())]
        gpatch = [gi + l for l in (
            "# successors from the device: sEnd / sStart hold the key of a start group (its smallest member) instead of a string",
            "start2NBPs = defaultdict(set); sStart = {}; sEnd = {}; _gid = {}",
            "for NBP in NBPs:",
            "    _S = self._b200_nbp_succ.get(NBP, ())",
            "    sEnd[NBP] = min(_S) if _S else None",
            "    for _N2 in _S: _gid[_N2] = sEnd[NBP]",
            "for NBP in NBPs:",
            "    sStart[NBP] = _gid.get(NBP)",
            "    if sStart[NBP] is not None: start2NBPs[sStart[NBP]].add(NBP)",
            "")]
        tail = tail[:g0] + gpatch + tail[g1:]
        # joined paths (`lib/Assembler.py:826`): their sequences come from `spb_path_sequences` on the device
        # instead of `reconstruct_sequence` in a fork pool
        old_call = "self.multimap(self.reconstruct_sequence, threads, newPaths, self.vertex2coords, self.ref2name, self.seqDict, self.ksize)"
        hits = [i for i, l in enumerate(tail) if old_call in l]
        assert len(hits) == 1, "the reference's run() changed: joined-path reconstruction not found"
        tail[hits[0]] = tail[hits[0]].replace(old_call, "self._b200_reconstruct(newPaths)")
    ns = dict(vars(mod))
    exec(compile("\n".join(lines[:start] + patch + tail), "<superpang run() with the B200 NBP collection>", "exec"), ns)
    return ns["run"]


def _seq_paths_nbps_from_device(i):
    """Replacement of `Assembler.get_seqPaths_NBPs` (`lib/Assembler.py:1121-1139`): same arguments through the
    reference's multiprocessing globals, same return value (the NBP-graph vertices whose NBP the sequence contains in
    its own orientation)."""
    names, seqPaths, psets, NBP2seq, vertex2NBP, seqDict = _ROW1["ref_class"].get_multiprocessing_globals()
    _ROW1["calls"] += 1
    mine = _ROW1["sets"].get(names[i], ())
    return {nv for pair in psets.values() for nv in pair if vertex2NBP[nv] in mine}


def run(b200, minlen, mincov, bubble_identity_threshold, genome_assignment_threshold, max_threads, debug=False, device_rows=True):
    """`Assembler.run(...)` of the reference (`lib/Assembler.py:260`) on top of a `superpang_b200.assembler.Assembler`."""
    Ref = reference_class()
    from graph_tool.all import Graph
    A = object.__new__(Ref)
    A.run = types.MethodType(patched_run(Ref, nbp_graph=device_rows), A)
    A.ksize, A.seqDict, A.ref2name, A.includedSeqs = b200.ksize, b200.seqDict, b200.ref2name, b200.includedSeqs
    A.vertex2coords = b200.vertex2coords
    A.seqPaths = dict(b200.seqPaths)
    A.seqLimits = b200.seqLimits
    A.seqLimitsPerSeq = b200.seqLimitsPerSeq
    A.DBG = Graph(directed=False)                                  # lib/Assembler.py:251-255
    A.DBG.add_edge_list(b200.edges)
    A.DBG.vp.comp = A.DBG.new_vertex_property("int16_t")
    A.DBG.ep.efilt = A.DBG.new_edge_property("bool")
    A._b200_components = b200.collect_nbps()
    if device_rows:
        _ROW1.update(sets=b200.seq_paths_nbps(), ref_class=Ref)
        A.get_seqPaths_NBPs = _seq_paths_nbps_from_device
        A._b200_nbp_succ = b200.nbp_successors()            # NBP graph (lib/Assembler.py:365-384) from spb_get_nbp_graph

        def _reconstruct(paths):
            _ROW1["joined"] = _ROW1.get("joined", 0) + len(paths)
            return b200.reconstruct_sequences([list(p) for p in paths]) if paths else []
        A._b200_reconstruct = _reconstruct
    try:
        return A.run(minlen, mincov, bubble_identity_threshold, genome_assignment_threshold, max_threads, debug)
    finally:
        _ROW1.update(sets=None)


def main(argv=None):
    """The reference CLI (`scripts/SuperPang.py`) with the B200 Assembler bound at its only call site (`:199`)."""
    from .assembler import Assembler
    from . import writers
    import superpang.scripts.SuperPang as cli_mod
    cli_mod.Assembler = Assembler
    cli_mod.write_fasta = writers.write_fasta          # NBPs.fasta / graph.fastg / *.core / *.aux through the native writer (:279-282)
    if argv is not None:
        sys.argv = [cli_mod.__file__] + list(argv)
    cli_mod.cli()


if __name__ == "__main__":
    main()
