"""TEST INFRASTRUCTURE ONLY — import the *unmodified* reference package from /root/reference.

The reference (ICHEC/qse 1.1.21) cannot be imported as-is in this image: it needs
``qutip`` and ``matplotlib`` (absent, no network) and package metadata for ``qse``
(``qse/__init__.py:38``).  This shim registers three inert stub modules and patches
``importlib.metadata.version`` for the name ``qse`` so that the reference's own pure-numpy
code (``qse.lattices``, ``qse.Qbits``, ``qse.Signal(s)``, ``qse.Cell``, ``qse.magnetic``,
``qse.calc.ExactSimulator``, ``qse.calc.blockade_radius``) runs unmodified.

Nothing under ``qse_b200/`` may import this module.  It is used by
``oracle/make_golden.py`` (fixture generation, this container only) and by the
container-only tests that are skipped when ``/root/reference`` does not exist (GPU box).
No reference source is copied; the package is imported from where it lies.
"""

from __future__ import annotations

import importlib
import importlib.metadata
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("QSE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "qse"))


class _Inert(types.ModuleType):
    """Module whose every attribute is a callable that refuses to run."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _missing(*a, **k):
            raise RuntimeError(f"{self.__name__}.{name} is a stub: not installed in this image")

        return _missing


def import_reference():
    """Return the reference ``qse`` package imported from REFERENCE_ROOT (with stubs)."""
    if "qse" in sys.modules and getattr(sys.modules["qse"], "__refshim__", False):
        return sys.modules["qse"]
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")

    for name in ("qutip", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
                 "matplotlib.patches", "matplotlib.cm", "matplotlib.collections"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Inert(name)
    # "import matplotlib.pyplot as plt" needs attribute access on the parent as well
    mpl = sys.modules["matplotlib"]
    for sub in ("pyplot", "colors", "patches", "cm", "collections"):
        if isinstance(mpl, _Inert):
            object.__setattr__(mpl, sub, sys.modules[f"matplotlib.{sub}"])

    real_version = importlib.metadata.version

    def _version(dist):
        if dist == "qse":
            return "1.1.21"  # pyproject.toml:16
        return real_version(dist)

    importlib.metadata.version = _version
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        qse = importlib.import_module("qse")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        importlib.metadata.version = real_version
    qse.__refshim__ = True
    return qse
