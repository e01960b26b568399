#!/usr/bin/env python3
"""
TEST INFRASTRUCTURE ONLY (build container: needs /root/reference) -- BASELINE config 1, the reference's own smoke test
(`scripts/test_SuperPang.py:23-28`, the "fast" pair of bundled genomes, `-i 1`), run with the UNMODIFIED reference CLI under
the import shims (`oracle/run_cli.py --impl reference`, `-b 1`: no mappy).  Freezes the sha256 of every output file the
north star names into tests/golden/cli_fast.json; the GPU tier runs the drop-in CLI on the same genomes and compares.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FILES = ("NBPs.fasta", "NBP2origins.tsv", "graph.fastg", "assembly.info.tsv", "assembly.fasta", "graph.NBP2origins.csv",
         "NBPs.core.fasta", "NBPs.accessory.fasta")
FAST = ("2636415981.fna", "2636416036.fna")
CLI_ARGS = ["--assume-complete", "-i", "1", "-b", "1", "--force-overwrite"]


def run_cli(impl, genomes, outdir, threads=8, extra=(), emu=False, timeout=3600):
    tmp = tempfile.mkdtemp(prefix="spb_cli_")
    cmd = [sys.executable, os.path.join(HERE, "run_cli.py"), "--impl", impl] + (["--emu"] if emu else []) + ["--", "-f"] + list(genomes) + \
        CLI_ARGS + ["-t", str(threads), "-o", outdir, "-d", tmp] + list(extra)
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    shutil.rmtree(tmp, ignore_errors=True)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
    return time.time() - t0


def digests(outdir):
    return {f: hashlib.sha256(open(os.path.join(outdir, f), "rb").read()).hexdigest() for f in FILES}


if __name__ == "__main__":
    td = "/root/reference/src/superpang/test_data"
    out = tempfile.mkdtemp(prefix="spb_cli_ref_")
    dt = run_cli("reference", [os.path.join(td, f) for f in FAST], out)
    d = dict(genomes=list(FAST), cli_args=CLI_ARGS, k=301, seconds_reference=round(dt, 1), cores=os.cpu_count(), sha256=digests(out),
             sizes={f: os.path.getsize(os.path.join(out, f)) for f in FILES})
    with open(os.path.join(ROOT, "tests", "golden", "cli_fast.json"), "w") as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d, indent=1))
    shutil.rmtree(out, ignore_errors=True)
