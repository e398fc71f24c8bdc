#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY.  Build recipe for `oracle/_ref/`.

Compiles the reference's own NLM kernel from the sources WHERE THEY LIE under
`/root/reference` (nothing is copied into the repository; `oracle/_ref/` is git-ignored):

  _filters      <- /root/reference/nd/_filters.pyx, UNMODIFIED.
                   (`cython -o oracle/_ref/_filters.c <pyx>`; gcc -O3, the flags of the
                   reference's setup.py:79-82.)
  _filters_aw   <- the same file streamed through the three-cast `sed` of SURVEY.md A.4
                   (`range(-f[i], ...)` -> `range(-<SIZE_TYPE>f[i], ...)`, nd/_filters.pyx:373-375),
                   i.e. the "as written" loops actually executing.  Secondary cross-check only.

  _change       <- /root/reference/nd/_change.pyx, UNMODIFIED (omnibus change detection).
                   Its only third-party arithmetic is `cython_gsl.gsl_cdf_chisq_P` (nd/_change.pyx:147-148); GSL is
                   not installed, so `cimport cython_gsl` resolves to the committed stand-in oracle/gsl_shim/
                   (a .pxd + a header with the chi-square CDF).  gcc -O3 -fopenmp (the reference's prange).

The reference's own build system (setup.py) is NOT run: it needs the whole package and
its `_filters.c` (Cython 0.29.13) does not compile on Python 3.12 (nd/_filters.c:216).
If /root/reference is absent (GPU box) the prebuilt files are used as they are.
"""
import os
import re
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/nd/_filters.pyx"


def _ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


SHIM = os.path.join(HERE, "gsl_shim")
SRC_CHANGE = "/root/reference/nd/_change.pyx"


def _compile(pyx_path, modname, cython_inc=(), extra=()):
    import numpy
    c_path = os.path.join(OUT, modname + ".c")
    so_path = os.path.join(OUT, modname + _ext_suffix())
    cmd = [sys.executable, "-m", "cython", "-X", "language_level=2", "-X", "emit_code_comments=False"]
    for d in cython_inc:
        cmd += ["-I", d]
    subprocess.check_call(cmd + ["-o", c_path, pyx_path])
    inc = sysconfig.get_paths()["include"]
    cc = ["gcc", "-O3", "-shared", "-fPIC", "-fno-strict-aliasing", "-w", "-I", inc, "-I", numpy.get_include()]
    for d in cython_inc:
        cc += ["-I", d]
    subprocess.check_call(cc + list(extra) + [c_path, "-o", so_path, "-lm"])
    return so_path


def change_built():
    return os.path.exists(os.path.join(OUT, "_change" + _ext_suffix()))


def build_change(force=False):
    """The reference's own omnibus change detection, compiled unmodified against the GSL stand-in."""
    if change_built() and not force:
        return True
    if not os.path.exists(SRC_CHANGE):
        return change_built()
    os.makedirs(OUT, exist_ok=True)
    _compile(SRC_CHANGE, "_change", cython_inc=(SHIM,), extra=("-fopenmp",))
    return True


def built():
    return all(os.path.exists(os.path.join(OUT, m + _ext_suffix()))
               for m in ("_filters", "_filters_aw"))


def build(force=False):
    """Build oracle/_ref; returns True if the compiled modules exist afterwards."""
    if built() and not force:
        build_change()
        return True
    if not os.path.exists(SRC):
        return built()
    os.makedirs(OUT, exist_ok=True)
    # (1) unmodified: compiled straight from the read-only tree (cython names the module
    #     after the file, so the source path can be used as it is).
    _compile(SRC, "_filters")
    # (2) as-written cross-check: a sed-equivalent stream edit, written only into _ref/.
    text = open(SRC).read()
    patched, n = re.subn(r"range\(-f\[([012])\], f\[([012])\] \+ 1\)",
                         r"range(-<SIZE_TYPE>f[\1], <SIZE_TYPE>f[\2] + 1)", text)
    assert n >= 3, "expected the three patch loops of nd/_filters.pyx:373-375"
    aw = os.path.join(OUT, "_filters_aw.pyx")
    with open(aw, "w") as fh:
        fh.write(patched)
    _compile(aw, "_filters_aw")
    with open(os.path.join(OUT, "__init__.py"), "w") as fh:
        fh.write("# built by oracle/build_ref.py from /root/reference/nd/_filters.pyx\n")
    build_change(force=force)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv) and build_change(force="--force" in sys.argv)
    print("oracle/_ref built:", ok, "| _change built:", change_built())
    sys.exit(0 if ok else 1)
