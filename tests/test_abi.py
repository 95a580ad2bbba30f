"""CPU: libhvp.so loads and exports every symbol include/hvp.h declares; API misuse is reported
through return codes; there is no CPU fallback (context creation fails without a device)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hvp.h")).read()
    return sorted(set(re.findall(r"\b(hvp_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    from hybrid_vehicle_platoon_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for nme in names:
        assert hasattr(L, nme), nme
    assert L.hvp_version() == 101


def test_no_cpu_fallback():
    import hybrid_vehicle_platoon_b200 as hvp
    if hvp._lib.lib().hvp_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hvp.Context(0)
    h = C.c_void_p()
    rc = hvp._lib.lib().hvp_ctx_create(0, C.byref(h))
    assert rc < 0 and "no CPU path" in hvp._lib.last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hybrid_vehicle_platoon_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "oracle" not in s.replace("the oracle", "").replace("against the oracle", "") or \
                    "import oracle" not in s and "from oracle" not in s and "hvp_oracle" not in s, f
