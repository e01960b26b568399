"""Worker for the world_size-2 CPU test of the multi-GPU path: gloo collectives + the host emulator build of the
kernels (tests/emu).  Usage: python mg_worker.py <rank> <world> <port> <case_json> <out_json>"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def gather_shards(eng, dist):
    """The sliced results of all ranks put together (before Engine.complete): the concatenation of the local parts must
    be exactly what one GPU computes."""
    import numpy as np
    got = {"counts": eng.counts.as_dict(), "origins": [a.copy() for a in eng.nbp_origins()]}
    for which in ("instance_vertices", "vertex2coords", "edges", "nbp_verts", "nbp_seqs"):
        t, first = eng.shard(which)
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, (first, t.cpu().numpy().copy()))
        if which != "edges":                     # contiguous, in rank order, no gaps
            pos = parts[0][0]
            width = 2 if which == "vertex2coords" else 1
            for first_r, a in parts:
                assert first_r == pos, (which, first_r, pos)
                pos += len(a) // width
        got[which] = np.concatenate([a for _, a in parts])
    return got


def check_shards(A, sliced):
    import numpy as np
    eng = A._eng
    c = eng.counts.as_dict()
    for k_ in ("n_vertices", "n_edges", "n_nbp", "nbp_vertices", "nbp_bases", "n_origin_pairs", "n_comp", "n_cycle_comp", "n_inits", "n_pieces", "n_junction"):
        assert sliced["counts"][k_] == c[k_], (k_, sliced["counts"][k_], c[k_])
    T = int(A._seq_off[-1])
    inst = eng.instance_vertices()[:T]
    got = sliced["instance_vertices"].view(np.uint32)[:T]
    assert np.array_equal(got, inst), "sliced per-position vertex ids differ"
    assert np.array_equal(sliced["vertex2coords"].view(np.uint32).reshape(-1, 2), eng.vertex2coords())
    assert np.array_equal(sliced["edges"].view(np.uint32).reshape(-1, 2), eng.edges())
    voff, verts, comp, cyc = eng.nbps()
    assert np.array_equal(sliced["nbp_verts"].view(np.uint32), verts)
    soff, seqs = eng.nbp_seqs()
    assert np.array_equal(sliced["nbp_seqs"], seqs)
    ooff, oidx = eng.nbp_origins()
    assert np.array_equal(sliced["origins"][0], ooff) and np.array_equal(sliced["origins"][1], oidx)


def main():
    rank, world, port, case_path, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from superpang_b200 import _capi
    from superpang_b200.assembler import Assembler
    import digests
    lib = _capi._declare(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libspb_emu.so")))
    case = json.load(open(case_path))
    out = []
    for c in case:
        if c.get("degenerate"):
            # orientation-ambiguous input: EVERY rank must refuse it (no rank may be left waiting in a collective)
            try:
                Assembler(dict(map(tuple, c["records"])), c["k"], engine=_capi.Engine(lib=lib), distributed=True).compact()
                out.append("accepted")
            except _capi.DegenerateInputError:
                out.append("degenerate")
            continue
        A = Assembler(dict(map(tuple, c["records"])), c["k"], engine=_capi.Engine(lib=lib), distributed=True)
        if os.environ.get("SPB_EXPECT_SLICED") == "1":
            assert getattr(A._eng, "mg_sliced", False), f"case went through {A._eng.mg_protocol}, not the sliced tail"
        shared = None
        if getattr(A._eng, "mg_sliced", False):
            # every rank copies the sequences of ITS NBPs to their place in one host buffer in POSIX shared memory (the e2e
            # path of bench.py at N > 1); compared with the complete result below
            shared = _capi.SharedHostBuffer(int(A._eng.counts.nbp_bases), tag=f"t{len(out)}")
            A._eng.shard_to_host("nbp_seqs", shared.tensor)
            dist.barrier()
        sliced = gather_shards(A._eng, dist) if getattr(A._eng, "mg_sliced", False) else None
        res = A.to_result()                      # (sliced tail: the first accessor recomputes everything on this rank)
        if sliced is not None:
            check_shards(A, sliced)
            import numpy as np
            n = int(A._eng.counts.nbp_bases)
            if shared.shared:
                assert np.array_equal(shared.tensor.numpy()[:n], A._eng.nbp_seqs()[1]), "shared host buffer differs from the complete NBP sequences"
            shared.close()
        out.append(digests.summarize(res))
    if rank == 0 or True:
        json.dump(out, open(f"{out_path}.{rank}", "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
