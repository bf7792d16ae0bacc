"""Import the UNMODIFIED reference (read-only at /root/reference) in this container.

Test infrastructure only: used by `make_golden.py` to generate the committed golden
vectors under tests/golden/.  The reference cannot travel to the GPU box, so nothing in
the `-m gpu` tests, `smoke()` or `bench.py` imports this module.

The reference package imports matplotlib/h5py/meshio/cmcrameri/gmsh at module scope
(src/pydrex/utils.py:9,13-16, src/pydrex/io.py:30-34); none are on its CPO hot path, so
they are replaced by inert stubs (recipe: SURVEY.md §8c).
"""

import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_SRC = "/root/reference/src"


def _stub(name):
    mod = types.ModuleType(name)
    mod.__getattr__ = lambda attr: MagicMock(name=f"{name}.{attr}")  # type: ignore
    mod.__path__ = []  # behave like a package
    return mod


def import_reference():
    """Return the reference `pydrex` package, importing it with stubs on first use."""
    if "pydrex" in sys.modules and getattr(sys.modules["pydrex"], "__file__", "").startswith(
        REFERENCE_SRC
    ):
        return sys.modules["pydrex"]
    if not os.path.isdir(REFERENCE_SRC):
        raise RuntimeError(f"reference not present at {REFERENCE_SRC}")
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    names = [
        "matplotlib", "matplotlib.axes", "matplotlib.projections", "matplotlib.pyplot",
        "matplotlib.collections", "matplotlib.legend_handler", "matplotlib.transforms",
        "matplotlib.colors", "matplotlib.cm", "matplotlib.figure", "matplotlib.tri",
        "h5py", "meshio", "cmcrameri", "cmcrameri.cm", "gmsh", "pyvista",
    ]
    for name in names:
        if name not in sys.modules:
            sys.modules[name] = _stub(name)
    mpl = sys.modules["matplotlib"]
    for name in names:
        if name.startswith("matplotlib."):
            setattr(mpl, name.split(".", 1)[1], sys.modules[name])

    class Axes:  # subclassed at src/pydrex/axes.py:12
        name = "stub"

    sys.modules["matplotlib.axes"].Axes = Axes
    pyplot = sys.modules["matplotlib.pyplot"]
    pyplot.rcParams = {"figure.figsize": (6.4, 4.8)}  # read at visualisation.py:18
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import pydrex  # noqa: E402

    return pydrex
