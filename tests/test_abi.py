"""The C-ABI library loads and exports every symbol include/conanmp.h declares (CPU, no compute)."""
import ctypes
import os
import re

import conan_fgw_b200 as cmp
from conftest import ROOT


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(cmp_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_builds_and_exports_every_declared_symbol():
    path = cmp.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # the ctypes table mirrors the header one to one
    assert set(cmp._lib.SIGNATURES) == decl
    assert cmp._lib.lib().cmp_version() >= 100
    assert cmp._lib.lib().cmp_last_error_string() is not None


def test_no_cpu_fallback():
    import pytest
    import torch

    m = cmp.SchNetNoSum(None, hidden_channels=16, num_filters=16, num_interactions=1, num_gaussians=8)
    b = cmp.synthetic.make_batch(1, 1, 5)
    with pytest.raises(cmp._lib.ConanMPError):
        m(b.z, b.pos, b.batch)          # CPU tensors are rejected, never silently computed
    with pytest.raises(cmp._lib.ConanMPError):
        cmp.radius_graph(torch.zeros(4, 3), 1.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "conan-fgw_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_import_alias_shares_submodules():
    """`conan_fgw_b200.<sub>` must be the very module objects the hyphen-named package uses (a second copy
    of nn / dp under the alias name breaks isinstance checks such as the weight pre-packing's)."""
    import sys

    import conan_fgw_b200 as pkg
    from conan_fgw_b200.dp import RegressionStep
    from conan_fgw_b200.nn import Linear
    from conan_fgw_b200.utils import to_dense_batch  # noqa: F401

    assert sys.modules["conan_fgw_b200.nn"] is sys.modules["conan-fgw_b200.nn"]
    assert sys.modules["conan_fgw_b200.dp"] is sys.modules["conan-fgw_b200.dp"]
    assert Linear is pkg.Linear and RegressionStep is pkg.dp.RegressionStep
