"""Crafted inputs that exercise the reference's edge cases (SURVEY.md §8a "behaviour confirmed ...").

Each case is (name, k, [(record name, sequence), ...]).  Deterministic (fixed seed)."""
import random

_COMP = str.maketrans("ACGT", "TGCA")


def rc(s):
    return s.translate(_COMP)[::-1]


def _rnd(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def crafted_cases(k=31, seed=3):
    rng = random.Random(seed)
    s = _rnd(rng, 600)
    snp = s[:300] + ("A" if s[300] != "A" else "C") + s[301:]
    hp = _rnd(rng, 100) + "A" * (2 * k) + _rnd(rng, 100)
    x, y, last = _rnd(rng, 200), _rnd(rng, 200), _rnd(rng, k)
    C = _rnd(rng, 500)
    D = _rnd(rng, 300)
    ins = s[:250] + _rnd(rng, 90) + s[250:]
    cases = [
        ("single", [("a", s)]),
        ("rc_duplicate", [("a", s), ("b", rc(s))]),
        ("snp_bubble", [("a", s), ("b", snp)]),
        ("short_records", [("a", s[:k]), ("b", s[:k + 1]), ("c", s[100:100 + k + 9])]),
        ("homopolymer_selfloop", [("a", hp)]),
        ("homopolymer_end", [("a", _rnd(rng, 80) + "A" * (k + 5))]),
        ("shared_last_kmer", [("a", x + last), ("b", y + last)]),
        ("shared_first_kmer", [("a", last + x), ("b", last + y)]),
        ("leading_zero_quirk", [("a", s), ("b", s[1:1 + k + 12])]),
        ("contained", [("a", s), ("b", s[100:400]), ("c", rc(s[200:500]))]),
        ("insertion", [("a", s), ("b", ins)]),
        ("pure_cycle", [("c", C + C[:k])]),
        ("cycle_and_subseq", [("c", C + C[:k]), ("d", C[100:300])]),
        ("cycle_rotations", [("c", C + C[:k]), ("d", C[250:] + C[:250 + k])]),
        ("cycle_rc", [("c", C + C[:k]), ("d", rc(C[250:] + C[:250 + k]))]),
        ("two_cycles", [("c", C + C[:k]), ("d", D + D[:k])]),
        ("three_components", [("a", s), ("b", _rnd(rng, 400)), ("c", C + C[:k])]),
        ("repeat_within", [("a", x + D + y + D + _rnd(rng, 150))]),
        ("inverted_repeat", [("a", x + D + y + rc(D) + _rnd(rng, 150))]),
        ("tips", [("a", s), ("b", s[:200] + _rnd(rng, 60)), ("c", _rnd(rng, 60) + s[400:])]),
        ("many_short", [(f"r{i}", s[i * 7:i * 7 + k + 1 + (i % 5)]) for i in range(40)]),
        ("nsplit_names", [("a_Nsplit_0", s[:200]), ("a_Nsplit_1", s[200:400]), ("b", s[50:300])]),
        # dinucleotide repeat of k + 1 bases: consecutive k-mers are reverse complements of each other (one vertex, a
        # self-loop reached with different attachment sides) -- found by the hypothesis test
        ("dinucleotide_selfloop", [("a", _rnd(rng, 90) + "CG" * ((k + 1) // 2) + _rnd(rng, 90)), ("b", s[:120])]),
    ]
    out = [(n, k, recs) for n, recs in cases]
    # a dinucleotide run ending a record: the pair (GAGAGAGAG, AGAGAGAGA) is walked a second time over the same
    # consecutive-id edge but on the other strand -> orientation-ambiguous, the reference asserts (found by hypothesis)
    out.append(("dinucleotide_tail", 9, [
        ("s0", "CGACAATAACTCGCGGGTTGACTCGCAGCTGTAACCCAGACCCGAGCAATACACTGCAAAA"),
        ("s3", "CGACAATAACTCGCGGGTTGACTCGCAGCTGTAACCCAGACCCGAGCAATACACTGCAAAAGTCGAGAAAGGCTGAAGTCCAATGTCTTGCTCTACACCGTCCTGCGAT"
               "CTGCAATGATCCGTCAGCGGTTAAATCAATAAGAGTTTGAGAGAGAGAG")]))
    return out


def write_fasta(path, recs):
    with open(path, "w") as f:
        for n, q in recs:
            f.write(f">{n}\n{q}\n")
