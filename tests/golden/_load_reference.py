"""Import the UNMODIFIED reference (read-only at /root/reference) with the three out-of-tree
shims SURVEY.md §8c lists.  Only the golden-vector generator uses this file; nothing that runs
on the GPU box may import it (/root/reference does not exist there).

Shims (no edits to /root/reference):
  1. numpy 2.x dropped ``numpy.core.numeric.Inf``      (shared_funcs.py:9)
  2. matplotlib is not installed                       (func_VAELE_DP_MQAM_shaping.py:12 ...)
  3. ``simulate_dispersion`` builds a ragged array that numpy >= 1.24 rejects (shared_funcs.py:49)
     -- only patched on request (the golden vectors feed their own deterministic tensors).
"""
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("VAEQ_REFERENCE_ROOT", "/root/reference")


def _install_shims():
    import numpy.core.numeric as _ncn  # noqa: deprecated alias, still importable
    if not hasattr(_ncn, "Inf"):
        _ncn.Inf = np.inf
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def load(module_name, subdir="optical_DP_channel"):
    """Import ``module_name`` from the reference tree (e.g. 'shared_funcs')."""
    _install_shims()
    path = os.path.join(REF_ROOT, subdir)
    if not os.path.isdir(path):
        raise FileNotFoundError(f"reference tree not found at {path}")
    # the reference modules import each other by bare name, so the directory must be on sys.path;
    # it is inserted at the FRONT so the reference's own shared_funcs wins over any drop-in.
    saved = list(sys.path)
    sys.path.insert(0, path)
    try:
        for name in ("shared_funcs", module_name):
            if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(REF_ROOT):
                del sys.modules[name]
        mod = importlib.import_module(module_name)
    finally:
        sys.path[:] = saved
    assert mod.__file__.startswith(REF_ROOT), mod.__file__
    return mod
