"""
Output writers of the path (writer half) on the native C-ABI writers `spb_write_fasta` /
`spb_write_origins_tsv`: same call signatures and byte-identical files as the reference's own code, so the CLI
(`scripts/SuperPang.py:279-294`) can bind them in place of `lib/utils.py:write_fasta` and its inline TSV loop
(INTEGRATION.md).  Sequences may be handed over as Python strings (reference form) or as a CSR that still lives on the
device (`write_fasta_csr`), which skips the Python string round trip for multi-gigabyte NBP sets.
"""
import numpy as np

from . import _capi


def write_fasta(seqDict, fasta, lib=None):
    """Drop-in for the reference's `write_fasta(seqDict, fasta)` (`lib/utils.py:62-65`): '>{name}\\n{seq}\\n' per item, in
    dict order."""
    names = list(seqDict)
    enc = [s.encode() if isinstance(s, str) else bytes(s) for s in seqDict.values()]
    off = np.zeros(len(enc) + 1, dtype=np.uint64)
    np.cumsum([len(e) for e in enc], out=off[1:])
    blob = np.frombuffer(b"".join(enc), dtype=np.uint8) if enc and off[-1] else np.zeros(1, dtype=np.uint8)
    _capi.write_fasta_native(fasta, names, blob, off, lib=lib)


def write_fasta_csr(fasta, names, seqs, seq_off, lib=None):
    """FASTA from a CSR (uint8 sequence bytes + uint64 offsets) in host or device memory."""
    _capi.write_fasta_native(fasta, names, seqs, seq_off, lib=lib)


def write_origins_table(path, first_column, node_names, node_origins, name2bin, n_genomes, lib=None):
    """The reference's node / edge origin tables (`scripts/SuperPang.py:284-294`): header
    '{first_column}\\tSourceContigs\\tSourceGenomes\\tNumberOfSourceGenomes', then per node its origins (source-sequence
    names, in the order given), their bins and the number of distinct bins over `n_genomes`.
    node_origins: per node an iterable of source names (what `contig.origins` holds); name2bin: source name -> bin."""
    orig_names = list(name2bin)
    index = {n: i for i, n in enumerate(orig_names)}
    bins = sorted(set(name2bin.values()))
    bindex = {b: i for i, b in enumerate(bins)}
    off = np.zeros(len(node_names) + 1, dtype=np.uint64)
    idx = []
    for i, origins in enumerate(node_origins):
        idx.extend(index[o] for o in origins)
        off[i + 1] = len(idx)
    idx = np.asarray(idx, dtype=np.uint32) if idx else np.zeros(1, dtype=np.uint32)
    header = f"{first_column}\tSourceContigs\tSourceGenomes\tNumberOfSourceGenomes\n"
    _capi.write_origins_tsv_native(path, header, list(node_names), off, idx, orig_names, [bindex[name2bin[n]] for n in orig_names], bins,
                                   n_genomes, lib=lib)
