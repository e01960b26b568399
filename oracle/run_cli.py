#!/usr/bin/env python3
"""
TEST INFRASTRUCTURE ONLY -- runs the reference's CLI (`scripts/SuperPang.py`, BASELINE config 1: the plumbing run) under the
import shims (graph_tool / mappy / speedict / mOTUlizer stand-ins; graph-tool, minimap2 and mOTUlizer are absent from this
image, so only `-i 1 -b 1` runs: no homogenize, no bubble alignment -- SURVEY 8(c)):

  --impl reference   the UNMODIFIED reference (staged copy, oracle/_ref/superpang)
  --impl dropin      the same CLI with `superpang_b200.assembler.Assembler` bound at `scripts/SuperPang.py:199`
                     (`superpang_b200.dropin.main`); --emu runs it on the kernel-logic emulator instead of the GPU

Everything after `--` goes to the CLI unchanged (-f, -k, -t, --lowmem, -o ...).  PYTHONHASHSEED is pinned because the CLI
joins Python sets of strings into its output lines.
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def setup_paths():
    sys.path.insert(0, HERE)
    import refrun
    root = refrun.reference_root()
    for p in (os.path.join(HERE, "shims"), root, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the CLI starts condense-edges.py as a child process, which must find the shims too
    os.environ["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "shims"), root, os.environ.get("PYTHONPATH", "")])
    return root


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["reference", "dropin"], required=True)
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest and a.rest[0] == "--" else a.rest
    if os.environ.get("PYTHONHASHSEED") != "0":
        os.environ["PYTHONHASHSEED"] = "0"
        os.execv(sys.executable, [sys.executable] + sys.argv)
    setup_paths()
    if a.impl == "reference":
        import superpang.scripts.SuperPang as cli_mod
        sys.argv = [cli_mod.__file__] + rest
        cli_mod.cli()
        return
    from superpang_b200 import dropin
    if a.emu:
        import ctypes
        from superpang_b200 import _capi
        from superpang_b200.assembler import Assembler
        lib = _capi._declare(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libspb_emu.so")))
        Assembler.engine_factory = staticmethod(lambda: _capi.Engine(lib=lib))
    dropin.main(rest)


if __name__ == "__main__":
    main()
