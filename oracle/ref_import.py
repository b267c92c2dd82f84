"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* upstream reference from /root/reference.

Only usable inside the dev container (the GPU box has no /root/reference).  It is used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by the
``-m "not gpu"`` tests that cross-check the numpy restatement against the real reference when
the reference tree happens to be present.  Nothing in the product package imports this.

The reference imports a few third-party modules that are not installed here (colorlog, kornia,
matplotlib, skimage, iopath, yacs).  None of them is on the arithmetic path we validate, so we
register empty stand-ins *before* importing (recipe from SURVEY.md section 8c).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MVAL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "triangulation.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install_stubs():
    try:
        import colorlog  # noqa: F401
    except Exception:
        _stub("colorlog", basicConfig=lambda **kw: None)
    try:
        import kornia  # noqa: F401
    except Exception:
        _stub("kornia")
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        m = _stub("matplotlib")
        p = _stub("matplotlib.pyplot")
        m.pyplot = p


def load_reference():
    """Returns (triangulation_module, evaluation_module, coreset_module) of the real reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    # The reference uses top-level ``from utils import ...``; our own package also has a ``utils``
    # sub-package but never as a top-level name, so putting the reference root first is safe.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("utils", "utils.triangulation", "utils.evaluation", "utils.coreset", "pose_estimators",
                 "pose_estimators.loss"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
    tri = importlib.import_module("utils.triangulation")
    ev = importlib.import_module("utils.evaluation")
    cs = importlib.import_module("utils.coreset")
    return tri, ev, cs
