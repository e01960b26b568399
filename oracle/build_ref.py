#!/usr/bin/env python3
"""
TEST INFRASTRUCTURE ONLY -- recipe that compiles the reference's own native helpers.

The reference's hot path uses three Cython modules (`src/superpang/lib/Compressor.pyx`,
`vtools.pyx`, `cutils.pyx`; built by the reference's `setup.py:5-7`).  This script compiles them
*from where they lie under /root/reference* (the .pyx files are staged in a temp dir because the
reference tree is read-only and cythonize writes the generated .c next to its input) and puts
the built extension modules into `oracle/_ref/superpang_lib/`.  `stage()` additionally places a runnable copy of the
reference package (its .py files, VERSION and the built extensions) under `oracle/_ref/superpang/`, so that
`bench.py --impl reference` and the drop-in CLI test can run the UNMODIFIED reference on the GPU box, where
/root/reference does not exist.  `oracle/_ref/` is a build output: git-ignored (no reference source enters the
repository history) but not gpurun-ignored, so it travels to the GPU box like the built .so files.

Usage:  python oracle/build_ref.py [--reference /root/reference]
"""
import argparse
import glob
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "superpang_lib")
MODULES = ("Compressor", "vtools", "cutils")


def built():
    return all(glob.glob(os.path.join(OUT, m + ".*.so")) for m in MODULES)


def build(reference="/root/reference", force=False):
    src = os.path.join(reference, "src", "superpang", "lib")
    if not os.path.isdir(src):
        return False
    if built() and not force:
        return True
    import numpy as np
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    inc = [sysconfig.get_paths()["include"], np.get_include()]
    with tempfile.TemporaryDirectory(prefix="spb_ref_") as tmp:
        for m in MODULES:
            shutil.copy(os.path.join(src, m + ".pyx"), os.path.join(tmp, m + ".pyx"))
            subprocess.check_call([sys.executable, "-m", "cython", "-3", os.path.join(tmp, m + ".pyx")],
                                  cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            cmd = ["gcc", "-O2", "-shared", "-fPIC", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
            for i in inc:
                cmd += ["-I", i]
            cmd += [os.path.join(tmp, m + ".c"), "-o", os.path.join(OUT, m + ext)]
            subprocess.check_call(cmd)
    return True


STAGED = os.path.join(HERE, "_ref", "superpang")


def staged():
    return os.path.exists(os.path.join(STAGED, "lib", "Assembler.py")) and all(
        glob.glob(os.path.join(STAGED, "lib", m + ".*.so")) for m in MODULES)


def stage(reference="/root/reference", force=False):
    """oracle/_ref/superpang/ = runnable copy of the reference package (build container only)."""
    src = os.path.join(reference, "src", "superpang")
    if staged() and not force:
        return True
    if not build(reference, force):
        return False
    for sub in ("", "lib", "scripts"):
        os.makedirs(os.path.join(STAGED, sub), exist_ok=True)
        for f in os.listdir(os.path.join(src, sub)):
            a = os.path.join(src, sub, f)
            if os.path.isfile(a) and not os.path.islink(a) and (f.endswith(".py") or f == "VERSION"):
                b = os.path.join(STAGED, sub, f)
                shutil.copyfile(a, b)
                os.chmod(b, 0o755 if sub == "scripts" else 0o644)
    for so in glob.glob(os.path.join(OUT, "*.so")):
        shutil.copyfile(so, os.path.join(STAGED, "lib", os.path.basename(so)))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = stage(a.reference, a.force)
    print("oracle/_ref built and staged" if ok else "reference not present; nothing built")
