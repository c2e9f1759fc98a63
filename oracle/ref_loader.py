"""Import the reference package built by oracle/build_ref.py (oracle/_ref/teHmm).

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference .py files are
Python 2; instead of editing them we install the shims listed in SURVEY.md
section 8c before import:

  builtins.xrange/reduce/unicode, np.float/np.int/np.alltrue aliases,
  collections.Iterable, sys.maxint, a stub `pybedtools` module,
  pkg_resources.parse_version, and the package dir on sys.path (py2 implicit
  relative import `from _track import runSum`, track.py:16).
"""
import builtins
import collections
import collections.abc
import functools
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(_HERE, "_ref")
REF_PKG = os.path.join(REF_ROOT, "teHmm")

_loaded = None


def available():
    return os.path.isdir(REF_PKG) and any(
        f.startswith("_hmm.") and f.endswith(".so") for f in os.listdir(REF_PKG))


def _install_shims():
    import numpy as np
    if not hasattr(builtins, "xrange"):
        builtins.xrange = range
    if not hasattr(builtins, "reduce"):
        builtins.reduce = functools.reduce
    if not hasattr(builtins, "unicode"):
        builtins.unicode = str
    for name, val in (("float", float), ("int", int), ("bool", bool),
                      ("alltrue", np.all)):
        if name not in np.__dict__:
            setattr(np, name, val)
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    if not hasattr(sys, "maxint"):
        sys.maxint = sys.maxsize
    if "pybedtools" not in sys.modules:
        stub = types.ModuleType("pybedtools")
        stub.__version__ = "0.0"
        stub.set_tempdir = lambda *a, **k: None
        stub.cleanup = lambda *a, **k: None
        stub.BedTool = object
        stub.Interval = object
        sys.modules["pybedtools"] = stub
    import warnings
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import pkg_resources  # noqa: F401
    except Exception:  # pragma: no cover
        stub = types.ModuleType("pkg_resources")
        stub.parse_version = lambda v: tuple(int(x) for x in v.split(".") if x.isdigit())
        sys.modules["pkg_resources"] = stub


def load():
    """Returns a namespace with the reference modules:
    .hmm .basehmm .emission .track .common ._hmm ._emission ._basehmm"""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError("oracle/_ref is not built; run `python oracle/build_ref.py` "
                          "in the container that has /root/reference")
    _install_shims()
    for p in (REF_ROOT, REF_PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    import logging
    ns = types.SimpleNamespace()
    # common.py:200-211 walks the root logger's handlers reading `.stream`; pytest installs handlers
    # without one.  Hide those for the duration of the import (the reference files stay unedited).
    root = logging.getLogger()
    hidden = [h for h in root.handlers if not hasattr(h, "stream")]
    for h in hidden:
        root.removeHandler(h)
    try:
        for m in ("common", "_hmm", "_basehmm", "track", "_emission", "basehmm",
                  "emission", "hmm"):
            setattr(ns, m, importlib.import_module("teHmm." + m))
    finally:
        for h in hidden:
            root.addHandler(h)
    _loaded = ns
    return ns
