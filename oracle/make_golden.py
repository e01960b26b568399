#!/usr/bin/env python3
"""
TEST INFRASTRUCTURE ONLY -- freezes outputs of the UNMODIFIED reference into `tests/golden/`.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

Writes
  tests/golden/crafted.json.gz  every case of tests/cases.py at k=31 and k=33: inputs, the complete raw NBP
                                list (vertices, sequence, origins, cycle flag), vertex2coords, edges,
                                seqPaths, seqLimits -- produced by oracle/refrun.py (reference code).
  tests/golden/synthetic.json   digests of reference runs on generator outputs (superpang_b200/synth.py),
                                small enough for the oracle and the GPU path to recompute anywhere.
  tests/golden/real_slice.fa.gz + real_slice.json
                                the first 120 kbp of two bundled genomes (reference test data) with the
                                reference's digests at k=301 -- a real-data fixture that travels.
The KAT digests of the full bundled genome sets (tests/golden/kat_A.json, kat_B.json) were produced with
`python oracle/refrun.py <fasta> 301 8` and match SURVEY.md §8(c).
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import digests  # noqa: E402
import refrun  # noqa: E402
import cases  # noqa: E402
from superpang_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def full(res):
    return dict(
        names=res["names"], n_inst=res["n_inst"], ncomp=res["ncomp"],
        vertex2coords=np.asarray(res["vertex2coords"]).tolist(),
        edges=np.asarray(res["edges"]).tolist(),
        seqPaths={n: np.asarray(v).tolist() for n, v in res["seqPaths"].items()},
        seqLimits=list(res["seqLimits"]),
        nbps=[[list(v), s, list(o), int(c), bool(cy)] for v, s, o, c, cy in res["nbps"]],
        orient_counts=[int(x) for x in res["orient_counts"]],
        orient_sets=res["orient_sets"],
        joined=[[list(p), s, list(o)] for p, s, o in res["joined"]],
        bubble_map=res["bubble_map"], bubble_oris=res["bubble_oris"],
    )


def main():
    tmp = "/tmp/spb_golden"
    os.makedirs(tmp, exist_ok=True)
    out = []
    for k in (31, 33):
        for name, kk, recs in cases.crafted_cases(k):
            fa = os.path.join(tmp, f"{name}_{k}.fa")
            cases.write_fasta(fa, recs)
            try:
                res = refrun.run_reference(fa, kk)
            except AssertionError:
                # the reference's own assertions fire (orientation-ambiguous input): parity is undefined there,
                # the new path must refuse (SPB_EDEGENERATE) instead of answering differently
                out.append(dict(name=name, k=kk, records=recs, degenerate=True))
                print("crafted", name, k, "reference asserts")
                continue
            out.append(dict(name=name, k=kk, records=recs, summary=digests.summarize(res), result=full(res)))
            print("crafted", name, k, out[-1]["summary"]["n_nbp"])
    with gzip.GzipFile(os.path.join(GOLD, "crafted.json.gz"), "wb", mtime=0) as g:
        g.write(json.dumps(dict(generator="oracle/make_golden.py via oracle/refrun.py (reference code)", cases=out)).encode())

    syn = []
    for (n, length, ani, seed, mode, k) in [(4, 20000, 0.97, 11, "snp", 31), (6, 30000, 0.96, 12, "block", 31),
                                            (6, 30000, 0.96, 12, "block", 101), (5, 60000, 0.95, 13, "block", 301),
                                            (3, 40000, 0.97, 14, "snp", 301), (4, 40000, 0.96, 15, "block", 501)]:
        names, seqs = synth.genome_set(n, length, ani, seed=seed, mode=mode)
        fa = os.path.join(tmp, f"syn_{seed}_{mode}.fa")
        synth.write_fasta(fa, names, seqs)
        res = refrun.run_reference(fa, k, threads=8)
        syn.append(dict(n_genomes=n, length=length, ani=ani, seed=seed, mode=mode, k=k, summary=digests.summarize(res)))
        print("synthetic", seed, mode, k, syn[-1]["summary"]["n_nbp"])
    json.dump(dict(generator="oracle/make_golden.py via oracle/refrun.py (reference code)", cases=syn),
              open(os.path.join(GOLD, "synthetic.json"), "w"), indent=1)

    td = os.path.join(refrun.REFERENCE, "src", "superpang", "test_data")
    recs = []
    for f in ("2636415981.fna", "2636416036.fna"):
        txt = open(os.path.join(td, f)).read().strip().lstrip(">")
        got = 0
        for rec in txt.split("\n>"):
            name, seq = rec.split("\n", 1)
            seq = "".join(seq.split())
            take = seq[:120000 - got]
            recs.append((name.split(" ")[0], take))
            got += len(take)
            if got >= 120000:
                break
    fa = os.path.join(tmp, "real_slice.fa")
    cases.write_fasta(fa, recs)
    with open(fa, "rb") as f, gzip.GzipFile(os.path.join(GOLD, "real_slice.fa.gz"), "wb", mtime=0) as g:
        g.write(f.read())
    res = refrun.run_reference(fa, 301, threads=8)
    json.dump(dict(generator="oracle/make_golden.py via oracle/refrun.py (reference code)", k=301,
                   input="first 120 kbp of 2636415981.fna and 2636416036.fna (reference test_data)",
                   summary=digests.summarize(res)), open(os.path.join(GOLD, "real_slice.json"), "w"), indent=1)
    print("real slice", digests.summarize(res)["n_nbp"])


if __name__ == "__main__":
    main()
