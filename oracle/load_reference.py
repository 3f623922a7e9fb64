"""ORACLE helper (test infrastructure): import the *unmodified* reference from /root/reference.

Works only where /root/reference exists (the build container); the GPU box never calls this.
Used by oracle/gen_golden.py to produce tests/golden/*.npz and by tests that are skipped when
the reference tree is absent.

Tricks:
* ``lfd/__init__.py`` imports matplotlib/sqlalchemy/tkinter (absent) -> register a stub ``lfd``
  package object whose ``__path__`` points at the reference so the root ``__init__`` never runs;
* ``fitsio`` is absent -> register ``lfd_b200.fitsio_lite`` under that name (covers the three calls
  the reference makes: detecttrails.py:113-114, removestars.py:96).
"""
import importlib
import importlib.util
import os
import sys
import types
import warnings

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lfd", "detecttrails"))


def load_processfield():
    """The reference's processfield module alone (needs only cv2 + numpy)."""
    path = os.path.join(REFERENCE_ROOT, "lfd", "detecttrails", "processfield.py")
    spec = importlib.util.spec_from_file_location("_ref_processfield", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_detecttrails():
    """The reference's ``lfd.detecttrails`` package (DetectTrails, process_field, remove_stars)."""
    if "lfd" not in sys.modules or not hasattr(sys.modules["lfd"], "__path__"):
        stub = types.ModuleType("lfd")
        stub.__path__ = [os.path.join(REFERENCE_ROOT, "lfd")]
        sys.modules["lfd"] = stub
    if "fitsio" not in sys.modules:
        here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if here not in sys.path:
            sys.path.insert(0, here)
        from lfd_b200 import fitsio_lite
        sys.modules["fitsio"] = fitsio_lite
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("lfd.detecttrails")
