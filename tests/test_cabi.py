"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/cpd_b200.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from probreg_b200 import _cabi


def _declared():
    txt = open(os.path.join(ROOT, "include", "cpd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cpd_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    lib = _cabi.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libcpd_b200.so does not export %s" % n
    assert sorted(_cabi.EXPORTED) == names      # the ctypes table covers the header, nothing more
    assert lib.cpd_version() >= 100


def test_no_cpu_fallback_without_a_gpu():
    if _cabi.lib().cpd_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_cabi.CpdError, match="no CPU path|no CUDA device"):
        _cabi.Handle(3)
    from probreg_b200 import cpd
    src = np.random.default_rng(0).random((10, 3))
    with pytest.raises(_cabi.CpdError):
        cpd.registration_cpd(src, src)
    with pytest.raises(_cabi.CpdError):
        cpd.RigidCPD(src).expectation_step(src, src, 0.1)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under probreg_b200/ may import, include or load it."""
    pkg = os.path.join(ROOT, "probreg_b200")
    bad = re.compile(r"^\s*(from|import)\s+\.*oracle|#include.*oracle|estep_oracle|c_oracle|cpd_oracle\s+import", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                body = open(os.path.join(dirpath, f)).read()
                assert not bad.search(body), "%s reaches into oracle/" % f


def test_argument_errors_mirror_the_reference():
    from probreg_b200 import cpd
    with pytest.raises(ValueError, match="Unknown transformation type"):     # probreg/cpd.py:454
        cpd.registration_cpd(np.zeros((3, 3)), np.zeros((3, 3)), tf_type_name="bogus")
    r = cpd.RigidCPD(np.zeros((4, 3)))
    with pytest.raises(AssertionError):                                       # probreg/cpd.py:73
        r.expectation_step(np.zeros(3), np.zeros((3, 3)), 1.0)
    assert cpd.EstepResult._fields == ("pt1", "p1", "px", "n_p")
    assert cpd.MstepResult._fields == ("transformation", "sigma2", "q")
