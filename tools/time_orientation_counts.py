#!/usr/bin/env python3
"""Times spb_get_nbp_orientation_counts at BASELINE config 3 scale; prints one JSON line per mode."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from superpang_b200 import _capi, synth  # noqa: E402
from superpang_b200.assembler import concat_sequences  # noqa: E402

for mode in sys.argv[1:] or ["snp", "block"]:
    names, seqs = synth.genome_set(100, 5_000_000, 0.96, seed=1, mode=mode)
    buf, off = concat_sequences(seqs)
    eng = _capi.Engine(device=0)
    eng.load(torch.from_numpy(buf).cuda(), off, 301)
    eng.build_dbg()
    c = eng.compact()
    eng.nbp_graph()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    counts = eng.nbp_orientation_counts()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps(dict(mode=mode, n_nbp=int(c["n_nbp"]), ms=round(ms, 2), counted=int(counts.sum()), max=int(counts.max()),
                          zero=int((counts == 0).sum()), one=int((counts == 1).sum()))))
