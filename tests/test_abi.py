"""The C-ABI library loads on a CPU-only box and exports exactly what include/slowflow_gpu.h declares.
No compute entry is called here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from slowflow_b200 import MTParams, VariationalParams, mt_params_default, variational_params_default
from slowflow_b200.api import ABI_SYMBOLS, load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "slowflow_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n.startswith(("sfgpu_", "sf_mt_", "variational", "epic"))))


def test_header_functions_are_exported(built):
    lib = load_library()
    decl = declared_functions()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), "libslowflow_gpu.so does not export %s" % name
    assert sorted(ABI_SYMBOLS) == decl


def test_defaults_match_reference_values(built):
    lib = load_library()
    p = VariationalParams()
    lib.variational_params_default(C.byref(p))
    q = variational_params_default()
    assert bytes(p) == bytes(q)
    assert (p.alpha, p.niter_outer, p.niter_inner, p.niter_solver) == (1.0, 5, 1, 30)
    assert abs(p.gamma - 0.71) < 1e-7 and abs(p.sor_omega - 1.9) < 1e-6 and p.delta == 0.0
    m = MTParams()
    lib.sf_mt_params_default(C.byref(m))
    assert bytes(m) == bytes(mt_params_default())


def test_no_cpu_fallback(built):
    """Without a CUDA device the handle API must refuse, loudly."""
    lib = load_library()
    if lib.sfgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.sfgpu_create(0, None, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.sfgpu_last_error()


def test_product_never_imports_oracle():
    """The shipped package and the CUDA sources must not reference oracle/."""
    bad = []
    for d in ("slowflow_b200", "include"):
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    txt = open(os.path.join(base, f), errors="ignore").read()
                    if re.search(r"pyoracle|sf_oracle|libsf_ref|libsf_oracle|from oracle|import oracle", txt):
                        bad.append(os.path.join(base, f))
    assert not bad, bad
