"""Import shim: just enough of the `gym` namespace for the reference's modules
to import in a container without gym.  Only used by
tests/golden/make_golden.py when generating fixtures; never shipped on a path
the product uses."""
from . import core, wrappers  # noqa: F401
from .core import Env, Wrapper  # noqa: F401


def make(*a, **k):
    raise RuntimeError("gym shim: no environments available")
