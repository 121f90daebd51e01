#!/usr/bin/env python
"""Recipe for ``baseline/_ref/``: a writable, importable copy of the UNMODIFIED reference for the
reference arm of ``bench.py`` and the drop-in tests on the GPU box (where ``/root/reference`` does not
exist).  ``baseline/_ref/`` is git-ignored (never part of this repository's history) but travels to the
GPU box with the gpurun snapshot.

    python baseline/make_ref.py [--src /root/reference] [--no-native]

What it does
  1. the contract's own command is tried first by hand and FAILS in this image
     (``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy>``: the
     ``fairseq.libnat`` torch C++ extension does not compile against torch 2.11 -- recorded in DESIGN.md);
  2. so the pure-Python packages the ST path needs are copied verbatim: ``fairseq/``, ``fairseq_cli/``,
     ``examples/speech_recognition/`` plus the top-level ``train.py`` / ``generate.py`` entry scripts;
  3. the two small native helpers those entry points import are compiled inside the copy:
     ``fairseq.data.data_utils_fast`` (Cython; ``batch_by_size``) and ``fairseq.libbleu`` (C++;
     imported by ``fairseq_cli/generate.py`` through ``fairseq/bleu.py``).  ``libnat`` /
     ``token_block_utils_fast`` (NAT and LM models) are not needed by the ST path and are skipped.

Nothing in the copy is edited: the run-time shims the old code needs under numpy 2 / Python 3.12 /
torch 2.11 are in-process monkeypatches (``oracle/ref_loader.py``).
"""
import argparse
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

SETUP_NATIVE = r'''
import numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
exts = [Extension("fairseq.libbleu",
                  sources=["fairseq/clib/libbleu/libbleu.cpp", "fairseq/clib/libbleu/module.cpp"],
                  extra_compile_args=["-std=c++11", "-O3"])]
exts += cythonize([Extension("fairseq.data.data_utils_fast", sources=["fairseq/data/data_utils_fast.pyx"],
                             language="c++", include_dirs=[numpy.get_include()],
                             extra_compile_args=["-std=c++11", "-O3"])], language_level=3)
setup(name="fairseq_native_helpers", ext_modules=exts, script_args=["build_ext", "--inplace"])
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("FBKST_REFERENCE_SRC", "/root/reference"))
    ap.add_argument("--no-native", action="store_true")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(args.src, "examples", "speech_recognition")):
        print("make_ref: no reference tree at %s (nothing to do)" % args.src)
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, "examples"))
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "build")
    for pkg in ("fairseq", "fairseq_cli"):
        shutil.copytree(os.path.join(args.src, pkg), os.path.join(DST, pkg), ignore=ignore)
    shutil.copytree(os.path.join(args.src, "examples", "speech_recognition"),
                    os.path.join(DST, "examples", "speech_recognition"), ignore=ignore)
    for f in ("train.py", "generate.py", "LICENSE"):
        shutil.copy(os.path.join(args.src, f), os.path.join(DST, f))
    if not args.no_native:
        with open(os.path.join(DST, "_setup_native.py"), "w") as f:
            f.write(SETUP_NATIVE)
        r = subprocess.run([sys.executable, "_setup_native.py"], cwd=DST, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            print(r.stdout[-3000:])
            print("make_ref: native helpers failed to build (train.py/generate.py entry points will not "
                  "import; the encoder-only reference arm does not need them)")
        shutil.rmtree(os.path.join(DST, "build"), ignore_errors=True)
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    print("make_ref: %d files under %s" % (n, DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
