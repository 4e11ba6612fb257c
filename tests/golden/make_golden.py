"""Generates tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libsf_ref.so = the unmodified
sources of /root/reference compiled in place).  Run here (needs /root/reference); the vectors travel.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from slowflow_b200 import ColorImage, Image  # noqa: E402
from oracle.pyoracle import Reference, SOR_LEX, SOR_REDBLACK  # noqa: E402
import helpers  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def two_frame(ref, w, h, name):
    im1, im2, wx0, wy0 = helpers.pair(w, h)
    out = {}
    for mode, tag in ((SOR_LEX, "lex"), (SOR_REDBLACK, "rb")):
        wx, wy = wx0.copy(), wy0.copy()
        ref.variational(wx, wy, im1, im2, None, mode)
        out["wx_" + tag], out["wy_" + tag] = wx.array.copy(), wy.array.copy()
    # operator level, straight from the reference's exported functions (variational_aux.h:12-29)
    L = ref.lib
    wim, mask = ColorImage(w, h), Image(w, h)
    L.image_warp(wim.ptr(), mask.ptr(), im2.ptr(), wx0.ptr(), wy0.ptr())
    D = [ColorImage(w, h) for _ in range(8)]
    L.get_derivatives(im1.ptr(), wim.ptr(), ref.deriv, *[d.ptr() for d in D])
    dps = L.compute_dpsis_weight(im1.ptr(), 5.0, ref.deriv)
    dpsis = Image(w, h)
    C.memmove(dpsis.buf.ctypes.data, dps.contents.data, dpsis.buf.nbytes)
    L.image_delete(dps)
    sh, sv = Image(w, h), Image(w, h)
    L.compute_smoothness(sh.ptr(), sv.ptr(), wx0.ptr(), wy0.ptr(), dpsis.ptr(), ref.deriv_flow, 0.5)
    du, dv = Image(w, h), Image(w, h)
    A = [Image(w, h) for _ in range(5)]
    L.compute_data_and_match(*[a.ptr() for a in A], mask.ptr(), du.ptr(), dv.ptr(), *[d.ptr() for d in D], 0.05,
                             0.71 * 0.5 / 3.0)
    out.update(warp=wim.array.copy(), mask=mask.array.copy(), dpsis=dpsis.array.copy(), sh=sh.array.copy(),
               sv=sv.array.copy(), ixx=D[3].array.copy(), ixy=D[4].array.copy(), iyz=D[7].array.copy(),
               a11=A[0].array.copy(), a12=A[1].array.copy(), a22=A[2].array.copy(), b1=A[3].array.copy(),
               b2=A[4].array.copy())
    L.sub_laplacian(A[3].ptr(), wx0.ptr(), sh.ptr(), sv.ptr())
    out["b1_lap"] = A[3].array.copy()
    L.sor_coupled(du.ptr(), dv.ptr(), *[a.ptr() for a in A], sh.ptr(), sv.ptr(), 7, 1.9)
    out["sor_du"], out["sor_dv"] = du.array.copy(), dv.array.copy()
    np.savez_compressed(os.path.join(OUT, name), w=w, h=h, **{k: v.astype(np.float32) for k, v in out.items()})
    print("wrote", name, {k: v.shape for k, v in out.items() if k.startswith("wx")})


if __name__ == "__main__":
    ref = Reference()
    two_frame(ref, 61, 45, "two_frame_61x45.npz")   # odd width -> stride 64 with 3 padding columns
    two_frame(ref, 96, 64, "two_frame_96x64.npz")
    if hasattr(ref.lib, "sf_ref_variational_mt"):
        import make_golden_mt
        make_golden_mt.main(ref)
