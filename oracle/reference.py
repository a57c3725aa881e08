"""Import the UNMODIFIED reference (``oracle/_ref``, see build_ref.py).  TEST INFRASTRUCTURE ONLY.

``load()`` puts the import stubs and ``oracle/_ref`` on ``sys.path`` and returns the ``permutect`` package.  When
``permutect_b200`` was imported first it has registered a two-line alias module ``permutect.parameters`` (so that ``.pt``
files pickle their hyper-parameters under the reference's class path); the alias is dropped in favour of the real
package, and the product's ``ModelParameters`` goes back to pickling under its own module path.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = os.path.join(HERE, "ref_stubs")
TARGET = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(TARGET, "permutect"))


def load():
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    for name in [n for n, m in list(sys.modules.items())
                 if (n == "permutect" or n.startswith("permutect.")) and getattr(m, "_permutect_b200_alias", False)]:
        del sys.modules[name]
    for path in (TARGET, STUBS):
        if path not in sys.path:
            sys.path.insert(0, path)
    importlib.invalidate_caches()
    pkg = importlib.import_module("permutect")
    assert os.path.abspath(os.path.dirname(pkg.__file__)).startswith(TARGET), pkg.__file__
    mine = sys.modules.get("permutect_b200.parameters")
    if mine is not None:
        mine.ModelParameters.__module__ = "permutect_b200.parameters"
    return pkg
