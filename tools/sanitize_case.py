#!/usr/bin/env python
"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck): odd geometries of the two-frame path, one
multi-frame window with a pyramid and occlusion reasoning, the operator twins."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import helpers  # noqa: E402
import mt_helpers as mh  # noqa: E402
from slowflow_b200 import ColorImage, Context, Image, variational_params_default  # noqa: E402

with Context(0) as ctx:
    for (w, h) in [(61, 45), (130, 67), (5, 5), (257, 33)]:
        im1, im2, wx, wy = helpers.pair(w, h)
        p = variational_params_default()
        p.niter_outer = 2
        ctx.variational(wx, wy, im1, im2, p)
        p.niter_inner, p.delta = 2, 0.5
        ctx.variational(wx, wy, im1, im2, p)
        assert np.isfinite(wx.array).all()
    ims, wx, wy = mh.window(97, 71, 3, zero_flow=True)
    q = mh.params(3, layers=2, niter_alter=2, niter_outer=2)
    g = mh.run_gpu(ctx, ims, wx, wy, q)
    assert np.isfinite(g["wx"].array).all()
    src = ColorImage.from_array((np.random.RandomState(0).rand(3, 45, 61) * 255).astype(np.float32))
    ctx.prescale(src, 0.6)
    ctx.get_derivatives(src, src)
    frames = [ColorImage.from_array(np.random.RandomState(k).rand(3, 40, 70).astype(np.float32) * 255) for k in range(4)]
    wxs, wys = [Image(70, 40) for _ in range(3)], [Image(70, 40) for _ in range(3)]
    for a in wxs + wys:
        a.buf[:] = 0
    ctx.variational_sequence(frames, wxs, wys, None)
    # round 2: integer-frame sequence (unpack kernel, resident ring), device min-cut, per-term data variant, EPIC
    from slowflow_b200 import synth  # noqa: E402
    w, h, n = 70, 50, 3
    raws = [np.ascontiguousarray(np.moveaxis(np.rint(synth.frame(w, h, t)).astype(np.uint8), 0, 2)) for t in range(n + 1)]
    u0, v0 = synth.initial_flow(w, h)
    xs, ys = [Image.from_array(u0) for _ in range(n)], [Image.from_array(v0) for _ in range(n)]
    p = variational_params_default()
    p.niter_outer, p.niter_solver = 2, 7
    ctx.variational_sequence_int(raws[:3], xs[:2], ys[:2], p)
    ctx.variational_sequence_int(raws[2:], xs[2:], ys[2:], p, continue_from_previous=True)
    import ctypes as C  # noqa: E402
    r = np.random.RandomState(1)
    d0, d1 = r.rand(37 * 23).astype(np.float32), r.rand(37 * 23).astype(np.float32)
    lab = np.zeros(37 * 23, np.int32)
    FP, IPi = C.POINTER(C.c_float), C.POINTER(C.c_int)
    assert ctx.lib.sfgpu_grid_mincut_dev(ctx.h, 37, 23, d0.ctypes.data_as(FP), d1.ctypes.data_as(FP), C.c_float(0.2), 0,
                                         lab.ctypes.data_as(IPi), None) == 0
    im, m, edges = synth.epic_case(97, 61, 120)
    fx, fy = Image(97, 61), Image(97, 61)
    ctx.epic(fx, fy, ColorImage.from_array(im), m, edges, None)
    assert np.isfinite(fx.array).all()
    # packed all-terms pass (two columns per thread): odd width with channel weights, un-normalised data term
    ims, wx, wy = mh.window(93, 41, 3)
    chw = ColorImage.from_array(np.random.RandomState(3).uniform(0.5, 1.5, size=(3, 41, 93)).astype(np.float32))
    for kw in (dict(robust_color=4, robust_color_eps=0.5), dict(dataterm=0, robust_color=3, robust_color_truncation=5.0)):
        g = mh.run_gpu(ctx, ims, wx, wy, mh.params(3, niter_alter=1, niter_outer=1, **kw), chw)
        assert np.isfinite(g["wx"].array).all()
os.environ["SLOWFLOW_GPU_MT_TERMS_SCALAR"] = "1"
with Context(0) as ctx:
    ims, wx, wy = mh.window(93, 41, 3)
    g = mh.run_gpu(ctx, ims, wx, wy, mh.params(3, niter_alter=1, niter_outer=1))
    assert np.isfinite(g["wx"].array).all()
os.environ["SLOWFLOW_GPU_MT_DATA_VARIANT"] = "1"
with Context(0) as ctx:
    ims, wx, wy = mh.window(97, 71, 3)
    g = mh.run_gpu(ctx, ims, wx, wy, mh.params(3, niter_alter=2, niter_outer=2))
    assert np.isfinite(g["wx"].array).all()
print("sanitize_case: ok")
