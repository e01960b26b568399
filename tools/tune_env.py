#!/usr/bin/env python3
"""Sweep library tuning knobs (environment variables read by libsuperpang_b200.so at every call) on ONE resident workload.

    python tools/tune_env.py --workload c3 --mode snp --out gpurun_out/tune.jsonl  "SPB_P1_SHAPE=0" "SPB_P1_SHAPE=1,SPB_PART_L1=8" ...

Each positional argument is one variant (comma-separated NAME=VALUE pairs; "-" = no variable).  Per variant: 2 warm-up steps,
3 timed steps (CUDA events, L2 flushed in between), the stage timers of the fastest step, and the device-side result digest
(all variants must agree).  A measurement tool: nothing here is on the product path.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--mode", default="snp")
    ap.add_argument("--k", type=int, default=301)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("variants", nargs="+")
    args = ap.parse_args()
    import torch
    import bench
    from superpang_b200 import _capi, digest
    bench.K = args.k
    w, names, buf, off, n_inst = bench.make_workload(args.workload, args.mode)
    dev = torch.device("cuda", 0)
    dbuf = torch.from_numpy(buf).to(dev)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    first = None
    for var in args.variants:
        env = dict(kv.split("=", 1) for kv in var.split(",") if "=" in kv)
        for k_, v in env.items():
            os.environ[k_] = v
        eng = _capi.Engine(device=0, stream=stream.cuda_stream)
        rec = dict(variant=var)
        try:
            times, stages = [], []
            for i in range(2 + args.steps):
                flush.fill_(1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.load(dbuf, off, args.k)
                eng.build_dbg()
                eng.compact()
                e1.record(stream)
                torch.cuda.synchronize()
                if i >= 2:
                    times.append(e0.elapsed_time(e1))
                    stages.append(dict(eng.stage_stats()))
            best = int(np.argmin(times))
            d = digest.result_digest(eng, graph=True)
            if first is None:
                first = d
            rec.update(ms_min=times[best], ms_median=float(np.median(times)), stages={k_: round(v, 3) for k_, v in stages[best].items()},
                       digest_equal=(d == first), gkmers_s=n_inst / times[best] / 1e6)
        except Exception as e:      # a variant that fails must not end the sweep
            rec.update(error=repr(e))
        eng.close()
        torch.cuda.empty_cache()
        for k_ in env:
            os.environ.pop(k_, None)
        line = json.dumps(rec)
        print(line, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")


if __name__ == "__main__":
    main()
