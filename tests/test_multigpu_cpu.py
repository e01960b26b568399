"""CPU tier: the multi-GPU orchestration (hash-partition + all-to-all + all-gather, superpang_b200/_capi.py
`build_dbg_distributed`) with world_size 2 and 4 over gloo, kernels executed by the host emulator.  Every rank
must end with exactly the reference's result."""
import json
import os
import socket
import subprocess
import sys

import pytest

import helpers

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return str(p)


@pytest.mark.parametrize("world,env", [
    (2, {}), (4, {}),
    (2, {"SPB_EMU_ORDER": "reverse"}),                               # displacement chains resolved at the owner
    (2, {"SPB_BUCKET_BITS": "3", "SPB_EMU_ORDER": "shuffle"}),       # owner sub-buckets
    (2, {"SPB_MG_PROTOCOL": "reps"}),                                # one representative per record travels back
    (2, {"SPB_MG_PROTOCOL": "links"}),                               # records travel, links travel back
    (2, {"SPB_MG_PROTOCOL": "owned"}),                               # owner computes: nothing but links travels
    (4, {"SPB_MG_PROTOCOL": "owned", "SPB_SMEM_AVG": "60", "SPB_EMU_ORDER": "reverse"}),
    (2, {"SPB_MG_PROTOCOL": "owned", "SPB_SMEM_AVG": "1", "SPB_EMU_ORDER": "shuffle"}),
    (8, {"SPB_MG_PROTOCOL": "owned", "SPB_SMEM_AVG": "200"}),          # the driver's largest scaling point
    (2, {"SPB_MG_PROTOCOL": "owned", "SPB_SMEM_LIMIT": "10"}),       # sub-buckets beyond the shared-memory table: tables in global memory
    (2, {"SPB_CHUNK_BITS": "8"}),                                    # protocol chosen by the duplicate-rate probe
    (2, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "60"}),          # record exchange on the partition kernels, all links to everybody
    (4, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "30", "SPB_EMU_ORDER": "reverse"}),
    (2, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "60", "SPB_MG_ALL_LINKS": "0", "SPB_EMU_ORDER": "shuffle"}),   # links to the slice owners, ids all-gathered
    (8, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "40", "SPB_MG_ALL_LINKS": "0"}),
    (2, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "2"}),           # regions overflow -> collective fallback to the exact path
    (2, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "60", "SPB_MG_ALL_LINKS": "2", "SPB_EXPECT_SLICED": "1"}),   # sliced tail forced on repeat-rich inputs too (remote vertices everywhere)
    (4, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "30", "SPB_MG_ALL_LINKS": "2", "SPB_EMU_ORDER": "shuffle", "SPB_EXPECT_SLICED": "1"}),
    (8, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "40", "SPB_MG_ALL_LINKS": "2", "SPB_EXPECT_SLICED": "1"}),
    (2, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "60", "SPB_MG_TAIL": "replicated"}),   # all links, graph + compaction on every rank
    (2, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "4", "SPB_PART_L2MAX": "4", "SPB_MG_ALL_LINKS": "2"}),   # level-2 partition in two passes (very large inputs)
], ids=lambda x: str(x) if isinstance(x, int) else (",".join(f"{k[4:].lower()}={v}" for k, v in x.items()) or "default"))
def test_distributed_build_matches_reference(emu_lib, tmp_path, world, env):
    crafted = [c for c in helpers.load_crafted() if c["k"] == 31 and not c.get("degenerate")]
    syn = helpers.load_json("synthetic.json")["cases"][:2]
    cases = [dict(records=c["records"], k=c["k"]) for c in crafted]
    want = [c["summary"] for c in crafted]
    for c in syn:
        sd = helpers.synth_seqdict(c["n_genomes"], c["length"], c["ani"], c["seed"], c["mode"])
        cases.append(dict(records=list(sd.items()), k=c["k"]))
        want.append(c["summary"])
    case_path = os.path.join(tmp_path, "cases.json")
    json.dump(cases, open(case_path, "w"))
    out = os.path.join(tmp_path, "out.json")
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "mg_worker.py"), str(r), str(world), port, case_path, out],
                              env=dict(os.environ, **env)) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    for r in range(world):
        got = json.load(open(f"{out}.{r}"))
        assert len(got) == len(want)
        for g, w in zip(got, want):
            helpers.assert_summary(g, w)


def test_sliced_tail_random_pangenomes(emu_lib, tmp_path):
    """Sliced graph build + compaction (4 ranks) on random pangenomes large enough that every rank owns positions, vertices
    and NBPs: every rank's complete() result must equal the single-engine result, and the concatenation of the local
    parts of all ranks must equal it too (checked inside the worker, tests/mg_worker.py: check_shards)."""
    from superpang_b200 import _capi
    from superpang_b200.assembler import Assembler
    import digests
    specs = [(6, 2500, 0.95, 101, "snp", 31), (5, 3000, 0.9, 102, "block", 31), (8, 1500, 0.97, 103, "snp", 21), (4, 4000, 0.85, 104, "block", 41),
             (3, 6000, 0.99, 105, "snp", 31)]
    cases, want = [], []
    for n, length, ani, seed, mode, k in specs:
        sd = helpers.synth_seqdict(n, length, ani, seed, mode)
        # a short circular-ish repeat and a sequence shorter than k, to cross the code paths of the crafted cases at this size
        sd["tiny"] = "ACGT" * 3
        cases.append(dict(records=list(sd.items()), k=k))
        A = Assembler(sd, k, engine=_capi.Engine(lib=emu_lib))
        want.append(digests.summarize(A.to_result()))
    case_path = os.path.join(tmp_path, "cases.json")
    json.dump(cases, open(case_path, "w"))
    out = os.path.join(tmp_path, "out.json")
    port = _free_port()
    env = {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "40", "SPB_MG_ALL_LINKS": "2", "SPB_EXPECT_SLICED": "1"}
    world = 4
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "mg_worker.py"), str(r), str(world), port, case_path, out],
                              env=dict(os.environ, **env)) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=900) == 0
    for r in range(world):
        got = json.load(open(f"{out}.{r}"))
        assert got == want


@pytest.mark.parametrize("world,env", [
    (2, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "60", "SPB_MG_ALL_LINKS": "2"}),          # sliced tail: the error flags are rank-local
    (4, {"SPB_MG_PROTOCOL": "part", "SPB_PART_AVG": "30", "SPB_MG_ALL_LINKS": "2"}),
    (2, {"SPB_MG_PROTOCOL": "owned"}),                                                       # replicated tail
], ids=lambda x: str(x) if isinstance(x, int) else ",".join(f"{k[4:].lower()}={v}" for k, v in x.items()))
def test_degenerate_input_is_refused_by_every_rank(emu_lib, tmp_path, world, env):
    """Orientation-ambiguous inputs (the reference asserts on them, `lib/Assembler.py:1088`, `:1104`) in between good ones:
    every rank raises DegenerateInputError for them -- also when only one rank's slice holds the offending positions -- and the
    ranks stay in step for the cases that follow."""
    crafted = helpers.load_crafted()
    deg = [c for c in crafted if c.get("degenerate")]
    good = [c for c in crafted if not c.get("degenerate") and c["k"] == 31][:2]
    assert deg
    cases = [dict(records=good[0]["records"], k=good[0]["k"])]
    for c in deg:
        cases.append(dict(records=c["records"], k=c["k"], degenerate=True))
    cases.append(dict(records=good[1]["records"], k=good[1]["k"]))
    case_path = os.path.join(tmp_path, "cases.json")
    json.dump(cases, open(case_path, "w"))
    out = os.path.join(tmp_path, "out.json")
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "mg_worker.py"), str(r), str(world), port, case_path, out],
                              env=dict(os.environ, **env)) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    for r in range(world):
        got = json.load(open(f"{out}.{r}"))
        assert got[1:1 + len(deg)] == ["degenerate"] * len(deg), got[1:1 + len(deg)]
        helpers.assert_summary(got[0], good[0]["summary"])
        helpers.assert_summary(got[-1], good[1]["summary"])
