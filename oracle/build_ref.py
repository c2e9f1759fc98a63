#!/usr/bin/env python
"""Build the UNMODIFIED reference (glennhickey/teHmm) hot path into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package (tehmm_b200/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it.

The reference is Python 2.7 + Cython.  Its five Cython modules compile from
their own sources with gcc, so we build them where they lie:

  1. read  /root/reference/{*.py,*.pyx}  (never written to, never committed);
  2. write patched copies ONLY under oracle/_ref/teHmm/ (git-ignored);
     the patch is the NumPy-2 alias fix from SURVEY.md section 8c:
        np.int_t -> np.int64_t ; dtype=np.int) -> dtype=np.int64)
        == np.float  -> == np.float64          (only in the .pyx files)
  3. cythonize with language_level=2 and build in place.

The .py files are copied verbatim; Python-2-isms are handled at import time by
oracle/ref_loader.py (builtins shims), not by editing the files.

oracle/_ref/ is listed in .gitignore (stays out of history) but NOT in
.gpurunignore, so the built .so files travel to the GPU box.
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("TEHMM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PKG = os.path.join(OUT, "teHmm")

PYX = ["_hmm.pyx", "_emission.pyx", "_basehmm.pyx", "_track.pyx", "_cfg.pyx"]
PY = ["__init__.py", "hmm.py", "basehmm.py", "emission.py", "track.py",
      "trackIO.py", "common.py", "modelIO.py", "cfg.py", "kmer.py"]


def patch_pyx(text):
    text = text.replace("np.int_t", "np.int64_t")
    text = text.replace("dtype=np.int)", "dtype=np.int64)")
    text = re.sub(r"== np\.float\b(?!\d)", "== np.float64", text)
    return text


def main():
    if not os.path.isdir(REF_SRC):
        print("reference sources not present at %s; keeping any prebuilt "
              "oracle/_ref as is" % REF_SRC)
        return 0 if os.path.isdir(PKG) else 1
    os.makedirs(PKG, exist_ok=True)
    for name in PY:
        src = os.path.join(REF_SRC, name)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(PKG, name))
    for name in PYX:
        with open(os.path.join(REF_SRC, name)) as f:
            text = f.read()
        with open(os.path.join(PKG, name), "w") as f:
            f.write(patch_pyx(text))

    import numpy as np
    from setuptools import setup, Extension
    from Cython.Build import cythonize

    exts = [Extension("teHmm." + n[:-4], [os.path.join(PKG, n)],
                      include_dirs=[np.get_include()],
                      extra_compile_args=["-O2", "-w"],
                      define_macros=[("NPY_NO_DEPRECATED_API", "0")])
            for n in PYX]
    cwd = os.getcwd()
    os.chdir(OUT)
    try:
        setup(name="teHmm_ref", script_args=["-q", "build_ext", "--inplace",
                                             "--build-temp", "build"],
              ext_modules=cythonize(exts, quiet=True,
                                    compiler_directives={"language_level": "2"}),
              packages=[])
    finally:
        os.chdir(cwd)
    shutil.rmtree(os.path.join(OUT, "build"), ignore_errors=True)
    built = [f for f in os.listdir(PKG) if f.endswith(".so")]
    print("built %d reference extension modules into %s" % (len(built), PKG))
    return 0 if len(built) == len(PYX) else 1


if __name__ == "__main__":
    sys.exit(main())
