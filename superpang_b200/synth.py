"""
Deterministic synthetic genome sets for benchmarks and property tests.

There is no network in the build/bench environment, so the BASELINE.json configurations
("10 genomes x 2 Mbp ANI 0.97", "100 genomes x 5 Mbp ANI 0.96", ...) are generated here:

  * ancestor: iid uniform ACGT, rejecting homopolymer / dinucleotide runs >= 50 (keeps the inputs
    away from the degenerate repeats on which the reference's own assertions fire,
    reference `lib/Assembler.py:1088`, `:1104-1106`);
  * mode "snp":   every base substituted independently with p = (1 - ANI) / 2, so pairwise identity
    is ~ANI (at k=301 almost every k-mer is unique: worst case for dedup volume);
  * mode "block": post-correction-like pangenome (the real inputs of the hot path are *corrected*
    genomes, reference `scripts/homogenize.py`): ancestor cut into 1-20 kbp blocks, 30 % of them
    accessory with a block-specific frequency, 0.5 % inverted, 1 % duplicated elsewhere, divergent
    islands of 150-2000 bp re-randomised with total fraction (1 - ANI); everything else identical.

Genome g uses `numpy.random.Generator(PCG64(seed + 1 + g))`; names follow the reference's
`bin|contig` convention (`scripts/SuperPang.py:137`).  Sequences are returned as uint8 ASCII arrays.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def _has_low_complexity(codes, run=50):
    """True if a homopolymer or dinucleotide repeat of length >= run exists (period <= 2)."""
    if codes.size < run:
        return False
    same = (codes[2:] == codes[:-2]).astype(np.int8)
    # longest run of ones in `same` >= run - 2
    d = np.diff(np.concatenate([[0], same, [0]]))
    starts, ends = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]
    return bool(starts.size and (ends - starts).max() >= run - 2)


def random_codes(rng, n):
    while True:
        c = rng.integers(0, 4, size=n, dtype=np.uint8)
        if not _has_low_complexity(c):
            return c


def revcomp_codes(c):
    return (3 - c)[::-1]


def _snp(rng, anc, ani):
    p = (1.0 - ani) / 2.0
    mask = rng.random(anc.size) < p
    out = anc.copy()
    out[mask] = (out[mask] + rng.integers(1, 4, size=int(mask.sum()), dtype=np.uint8)) & 3
    return out


def _block_layout(seed, length):
    rng = np.random.Generator(np.random.PCG64(seed))
    cuts = [0]
    while cuts[-1] < length:
        cuts.append(cuts[-1] + int(rng.integers(1000, 20001)))
    cuts[-1] = length
    cuts = np.asarray(cuts)
    nb = cuts.size - 1
    accessory = rng.random(nb) < 0.30
    freq = rng.uniform(0.1, 0.9, size=nb)
    return cuts, accessory, freq


def _block(rng, anc, ani, layout):
    cuts, accessory, freq = layout
    nb = cuts.size - 1
    present = ~accessory | (rng.random(nb) < freq)
    inverted = rng.random(nb) < 0.005
    dup = np.nonzero(present & (rng.random(nb) < 0.01))[0]
    order = [b for b in range(nb) if present[b]]
    for b in dup:                                  # duplicated elsewhere
        order.insert(int(rng.integers(0, len(order) + 1)), int(b))
    parts = []
    for b in order:
        seg = anc[cuts[b]:cuts[b + 1]]
        parts.append(revcomp_codes(seg) if inverted[b] else seg)
    g = np.concatenate(parts) if parts else anc[:0].copy()
    # divergent islands, total fraction (1 - ani)
    target = int((1.0 - ani) * g.size)
    done = 0
    while done < target and g.size > 4000:
        ln = int(rng.integers(150, 2001))
        st = int(rng.integers(0, g.size - ln))
        g[st:st + ln] = rng.integers(0, 4, size=ln, dtype=np.uint8)
        done += ln
    return g


def genome_set(n_genomes, length, ani=0.97, seed=1, mode="snp"):
    """Return (names, [uint8 ASCII arrays]) -- one contig per genome."""
    rng0 = np.random.Generator(np.random.PCG64(seed))
    anc = random_codes(rng0, length)
    layout = _block_layout(seed + 7919, length) if mode == "block" else None
    names, seqs = [], []
    for g in range(n_genomes):
        rng = np.random.Generator(np.random.PCG64(seed + 1 + g))
        if mode == "snp":
            c = _snp(rng, anc, ani)
        elif mode == "block":
            c = _block(rng, anc, ani, layout)
        else:
            raise ValueError(mode)
        if _has_low_complexity(c):
            c = _snp(rng, c, 0.999)
        names.append(f"g{g:04d}|c0")
        seqs.append(_ACGT[c])
    return names, seqs


def write_fasta(path, names, seqs, width=0):
    with open(path, "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + n.encode() + b"\n")
            f.write(s.tobytes() if isinstance(s, np.ndarray) else s.encode())
            f.write(b"\n")


def n_instances(seqs, k):
    return int(sum(max(0, len(s) - k + 1) for s in seqs if len(s) > k))


def genome_set_device(n_genomes, length, ani, seed, device):
    """SNP-mode genome set generated ON THE DEVICE (torch RNG; for workloads too large to generate and ship from the host
    within a benchmark run, e.g. BASELINE config 4 at 4e9 bases): ancestor iid uniform, every base of genome g substituted
    independently with p = (1 - ani) / 2.  Deterministic for a given (seed, GPU model): every rank of a multi-GPU run
    generates the same bytes.  Returns (uint8 ASCII device tensor of n_genomes * length bases, uint64 offsets)."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    anc = torch.randint(0, 4, (length,), dtype=torch.uint8, device=device, generator=gen)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    out = torch.empty(n_genomes * length, dtype=torch.uint8, device=device)
    p = (1.0 - ani) / 2.0
    for g in range(n_genomes):
        gen.manual_seed(int(seed) + 1 + g)
        mask = torch.rand(length, device=device, generator=gen) < p
        shift = torch.randint(1, 4, (length,), dtype=torch.uint8, device=device, generator=gen)
        c = torch.where(mask, (anc + shift) & 3, anc)
        out[g * length:(g + 1) * length] = lut[c.long()]
    off = np.arange(n_genomes + 1, dtype=np.uint64) * np.uint64(length)
    return out, off
