"""The UNMODIFIED reference for golden-vector generation and differential oracle tests: a thin alias of
oracle/reference.py (loader + shims).  The differential tests run only where the reference SOURCE tree exists (the
build container); nothing on the `-m gpu` path imports this module."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference as _reference  # noqa: E402

REFERENCE_ROOT = _reference.SOURCE_ROOT


def available():
    """Only the source tree counts here: the differential tests are a build-container check."""
    return os.path.isdir(os.path.join(_reference.SOURCE_ROOT, "mass"))


load = _reference.load
