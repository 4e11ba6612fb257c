"""CPU checks of the EPIC checker itself (oracle/_ref = the reference's epic.cpp / epic_aux.cpp + ref_glue): no GPU needed.

The one foreign piece on the checker's side is the stand-in for LAPACK's sgels_ (oracle/ref_glue/lapack_stub.c; LAPACK is
not installed in this image).  It is pinned here against numpy's least-squares solver on systems of the exact shape
fit_localaffine builds (epic_aux.cpp:438-481), and the whole reference epic() is run once on synthetic matches."""
import ctypes as C

import numpy as np
import pytest

from slowflow_b200 import ColorImage, Image, synth
from slowflow_b200.api import EpicParams, epic_params_default
from slowflow_b200.image import color_image_t, image_t

pytestmark = pytest.mark.usefixtures("built")


def test_sgels_stand_in_matches_numpy_lstsq(reference):
    L = reference.lib
    FP, IPi = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.sgels_.argtypes = [C.c_char_p, IPi, IPi, IPi, FP, IPi, FP, IPi, FP, IPi, IPi]
    r = np.random.RandomState(0)
    for nn in (5, 26, 100):
        n = 2 * (nn + 4)
        # rows of the reference's system: (x c, y c, 0, 0, c, 0 | (x + wx) c) and (0, 0, x c, y c, 0, c | (y + wy) c)
        x, y, c = r.uniform(0, 1000, nn + 4), r.uniform(0, 400, nn + 4), r.uniform(1e-3, 1.0, nn + 4)
        wx, wy = 0.01 * x - 0.02 * y + 3 + r.normal(0, 0.2, nn + 4), 0.015 * y + 1 + r.normal(0, 0.2, nn + 4)
        A = np.zeros((n, 6), np.float32)
        b = np.zeros(n, np.float32)
        A[0::2, 0], A[0::2, 1], A[0::2, 4] = x * c, y * c, c
        A[1::2, 2], A[1::2, 3], A[1::2, 5] = x * c, y * c, c
        b[0::2], b[1::2] = (x + wx) * c, (y + wy) * c
        ref = np.linalg.lstsq(A.astype(np.float64), b.astype(np.float64), rcond=None)[0]
        a_cm = np.ascontiguousarray(A)  # row-major (n, 6) == column-major 6 x n with lda = 6: one equation per column
        rhs = b.copy()
        m, nrhs, lda, ldb, info = C.c_int(6), C.c_int(1), C.c_int(6), C.c_int(n), C.c_int(0)
        nc = C.c_int(n)
        wq, lwork = C.c_float(0), C.c_int(-1)
        L.sgels_(b"Transposed", C.byref(m), C.byref(nc), C.byref(nrhs), a_cm.ctypes.data_as(FP), C.byref(lda), rhs.ctypes.data_as(FP),
                 C.byref(ldb), C.byref(wq), C.byref(lwork), C.byref(info))
        work = np.zeros(max(1, int(wq.value)), np.float32)
        lwork = C.c_int(work.size)
        L.sgels_(b"Transposed", C.byref(m), C.byref(nc), C.byref(nrhs), a_cm.ctypes.data_as(FP), C.byref(lda), rhs.ctypes.data_as(FP),
                 C.byref(ldb), work.ctypes.data_as(FP), C.byref(lwork), C.byref(info))
        assert info.value == 0
        assert np.allclose(rhs[:6], ref, rtol=2e-3, atol=2e-3), (nn, rhs[:6], ref)


def test_reference_epic_runs_and_interpolates(reference):
    L = reference.lib
    IP, CP = C.POINTER(image_t), C.POINTER(color_image_t)
    L.sf_ref_epic.argtypes = [IP, IP, CP, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(EpicParams)]
    L.sf_ref_epic.restype = None
    w, h, n = 200, 120, 400
    im, m, edges = synth.epic_case(w, h, n)
    ci = ColorImage.from_array(im)
    u, v = synth.gt_flow(w, h)
    for method in (b"LA", b"NW"):
        p = epic_params_default()
        p.method = method
        fx, fy, e = Image(w, h), Image(w, h), edges.copy()
        L.sf_ref_epic(fx.ptr(), fy.ptr(), ci.ptr(), m.ctypes.data, n, 4, e.ctypes.data, C.byref(p))
        assert np.allclose(e, edges + np.float32(0.001))  # euc added to the caller's array
        epe = float(np.sqrt((fx.array - u) ** 2 + (fy.array - v) ** 2).mean())
        assert np.isfinite(fx.array).all() and epe < (0.5 if method == b"LA" else 1.5), (method, epe)
