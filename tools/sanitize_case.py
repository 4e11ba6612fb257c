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
print("sanitize_case: done")
