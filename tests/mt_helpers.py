"""Helpers for the multi-frame parity tests."""
import ctypes as C

import numpy as np

from slowflow_b200 import ColorImage, Image, mt_params_default, synth
from slowflow_b200.image import color_image_t
from slowflow_b200.params import MTParams

CP = C.POINTER(color_image_t)


def window(w, h, S=3, seed=20170721, zero_flow=False, noise=0.25):
    frames, wx, wy = synth.window_case(w, h, S, seed, noise, zero_flow)
    return [ColorImage.from_array(f) for f in frames], Image.from_array(wx), Image.from_array(wy)


def params(S=3, **kw):
    p = mt_params_default()
    p.S = S
    p.hbit = 0  # synthetic intensities are 0..255
    for k, v in kw.items():
        if k in ("rho", "omega"):
            for a, x in enumerate(v):
                getattr(p, k)[a] = x
        else:
            setattr(p, k, v)
    return p


def clone_params(p):
    return MTParams.from_buffer_copy(bytes(p))


def frame_array(ims):
    return (CP * len(ims))(*[C.pointer(f.c) for f in ims])


def run_cpu(lib, prefix, ims, wx0, wy0, p, sor_mode, channel_w=None):
    """normalize + Variational_MT::variational through `sf_ref_*` (reference objects) or `sfo_*` (restatement)."""
    ims = [f.copy() for f in ims]
    arr = frame_array(ims)
    q = clone_params(p)
    getattr(lib, prefix + "normalize")(arr, len(ims), C.byref(q))
    wx, wy, occ = wx0.copy(), wy0.copy(), Image(wx0.width, wx0.height)
    avg, st = (C.c_float * 2)(), (C.c_int * 2)()
    getattr(lib, prefix + "variational_mt")(wx.ptr(), wy.ptr(), arr, C.byref(q), channel_w.ptr() if channel_w else None,
                                            occ.ptr(), avg, sor_mode, st)
    return dict(wx=wx, wy=wy, occ=occ, avg=(avg[0], avg[1]), stats=list(st), params=q, frames=ims)


def run_gpu(ctx, ims, wx0, wy0, p, channel_w=None):
    ims = [f.copy() for f in ims]
    q = clone_params(p)
    ctx.normalize(ims, q)
    wx, wy, occ = wx0.copy(), wy0.copy(), Image(wx0.width, wx0.height)
    avg = ctx.variational_mt(wx, wy, ims, q, channel_w, occ)
    st = ctx.mt_stats()
    return dict(wx=wx, wy=wy, occ=occ, avg=avg, stats=st, params=q, frames=ims)
