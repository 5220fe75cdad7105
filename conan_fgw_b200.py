"""Import alias: ``import conan_fgw_b200`` -> the package in ``conan-fgw_b200/``.

The package directory keeps the reference's hyphenated name; Python identifiers cannot contain a
hyphen, so this one-liner module loads it by path and re-exports it under an importable name.
"""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("conan-fgw_b200")
# alias the package AND every submodule it has loaded: `from conan_fgw_b200.dp import RegressionStep` must return the
# classes the package itself uses (a second copy of nn / dp under the alias name would break isinstance checks)
sys.modules[__name__] = _pkg
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("conan-fgw_b200."):
        sys.modules[__name__ + _name[len("conan-fgw_b200"):]] = _mod
