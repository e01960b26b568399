"""GPU worker (run under torchrun, one rank per GPU): multi-GPU DBG build over NCCL with the sm_100a library,
checked against the reference goldens on every rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import digests
    import helpers
    from superpang_b200 import _capi
    from superpang_b200.assembler import Assembler
    cases = []
    for c in helpers.load_json("synthetic.json")["cases"]:
        cases.append((helpers.synth_seqdict(c["n_genomes"], c["length"], c["ani"], c["seed"], c["mode"]), c["k"], c["summary"]))
    rs = helpers.load_json("real_slice.json")
    cases.append((helpers.real_slice_records(), rs["k"], rs["summary"]))
    for c in [c for c in helpers.load_crafted() if not c.get("degenerate")][:22]:
        cases.append((dict(map(tuple, c["records"])), c["k"], c["summary"]))
    for fa, name in (("A.fa", "kat_A.json"), ("B.fa", "kat_B.json")):
        path = os.path.join(helpers.DATA, fa)
        if os.path.exists(path):
            cases.append((path, 301, helpers.load_json(name)))
    used = set()
    from mg_worker import gather_shards, check_shards
    n_sliced = 0
    for proto in (None, "part", "part-all", "part-slice", "owned", "links", "reps"):           # None: chosen by the duplicate-rate probe
        if proto:
            os.environ["SPB_MG_PROTOCOL"] = proto.split("-")[0]
            # part-all: the sliced tail (graph build + compaction partitioned over the ranks) on every case, repeat-rich or not
            os.environ["SPB_MG_ALL_LINKS"] = "0" if proto.endswith("-slice") else "2" if proto.endswith("-all") else "1"
            if proto.startswith("part"):
                os.environ["SPB_PART_AVG"] = "300"           # both partition levels on the small cases
            else:
                os.environ.pop("SPB_PART_AVG", None)
        for sd, k, want in cases:
            if proto and isinstance(sd, str):
                continue                                      # the large KAT sets run once, with the probe's choice
            A = Assembler(sd, k, engine=_capi.Engine(device=local, stream=torch.cuda.current_stream().cuda_stream), distributed=True)
            assert A._eng.kind == "cuda sm_100a"
            used.add(A._eng.mg_protocol)
            sliced = gather_shards(A._eng, dist) if getattr(A._eng, "mg_sliced", False) else None
            assert sliced is not None or proto != "part-all"
            res = A.to_result()                   # (sliced tail: the first accessor recomputes everything on this rank)
            if sliced is not None:
                check_shards(A, sliced)           # concatenation of the local parts of all ranks == the complete result
                n_sliced += 1
            helpers.assert_summary(digests.summarize(res), want)
    assert used >= {"part", "owned", "links", "reps"}, used
    assert n_sliced >= len(cases) - 2, n_sliced
    dist.barrier()
    if dist.get_rank() == 0:
        print(f"multi-GPU parity OK on {dist.get_world_size()} ranks, {len(cases)} cases")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
