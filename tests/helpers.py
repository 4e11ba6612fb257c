"""Shared helpers for the parity tests: seeded inputs and oracle call chains."""
import ctypes as C

import numpy as np

from slowflow_b200 import ColorImage, Image, synth


def rng_plane(w, h, seed, lo=-1.0, hi=1.0):
    r = np.random.RandomState(seed)
    return Image.from_array(r.uniform(lo, hi, size=(h, w)).astype(np.float32))


def pair(w, h, seed=20170721, noise=0.25):
    im1, im2, wx, wy = synth.two_frame_case(w, h, seed, noise)
    return ColorImage.from_array(im1), ColorImage.from_array(im2), Image.from_array(wx), Image.from_array(wy)


def new_like(im):
    return Image(im.width, im.height)


def new_color_like(im):
    return ColorImage(im.width, im.height)


def oracle_system(oracle, im1, im2, wx, wy, du=None, dv=None, half_alpha=0.5, hd=0.0, hg=0.71 * 0.5 / 3.0):
    """warp -> derivatives -> smoothness -> data term -> laplacian with the oracle (variational.c:40-55)."""
    L = oracle.lib
    w, h = wx.width, wx.height
    wim, mask = ColorImage(w, h), Image(w, h)
    L.sfo_image_warp(wim.ptr(), mask.ptr(), im2.ptr(), wx.ptr(), wy.ptr(), 1)
    D = [ColorImage(w, h) for _ in range(8)]
    L.sfo_get_derivatives(im1.ptr(), wim.ptr(), *[d.ptr() for d in D])
    du = du or Image(w, h)
    dv = dv or Image(w, h)
    dps = L.sfo_compute_dpsis_weight(im1.ptr(), 5.0)
    dpsis = Image(w, h)
    C.memmove(dpsis.buf.ctypes.data, dps.contents.data, dpsis.buf.nbytes)
    L.sfo_image_delete(dps)
    uu, vv = Image(w, h), Image(w, h)
    uu.buf[:] = wx.buf + du.buf
    vv.buf[:] = wy.buf + dv.buf
    sh, sv = Image(w, h), Image(w, h)
    L.sfo_compute_smoothness(sh.ptr(), sv.ptr(), uu.ptr(), vv.ptr(), dpsis.ptr(), half_alpha)
    A = [Image(w, h) for _ in range(5)]
    L.sfo_compute_data_and_match(*[a.ptr() for a in A], mask.ptr(), du.ptr(), dv.ptr(), *[d.ptr() for d in D], hd, hg)
    raw = [a.copy() for a in A]
    L.sfo_sub_laplacian(A[3].ptr(), wx.ptr(), sh.ptr(), sv.ptr())
    L.sfo_sub_laplacian(A[4].ptr(), wy.ptr(), sh.ptr(), sv.ptr())
    return dict(wim=wim, mask=mask, D=D, dpsis=dpsis, sh=sh, sv=sv, A=A, raw=raw, du=du, dv=dv)


def close(a, b, atol, rtol=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= atol + rtol * np.abs(b)))


def maxdiff(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
