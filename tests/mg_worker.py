"""Worker for the world_size-2 CPU test of the multi-GPU path: gloo collectives + the host emulator build of the
kernels (tests/emu).  Usage: python mg_worker.py <rank> <world> <port> <case_json> <out_json>"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


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
        A = Assembler(dict(map(tuple, c["records"])), c["k"], engine=_capi.Engine(lib=lib), distributed=True)
        out.append(digests.summarize(A.to_result()))
    if rank == 0 or True:
        json.dump(out, open(f"{out_path}.{rank}", "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
