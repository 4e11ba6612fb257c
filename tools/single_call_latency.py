#!/usr/bin/env python
"""Latency of ONE drop-in call sfgpu_variational() at 2560x1440 (what an unmodified adaptiveFR.cpp:574 loop sees):
pageable caller buffers vs buffers page-locked once with sfgpu_host_register."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slowflow_b200 import ColorImage, Context, Image, synth  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2560, 1440)
im1, im2, wx0, wy0 = synth.two_frame_case(W, H)
a, b = ColorImage.from_array(im1), ColorImage.from_array(im2)
with Context(0) as ctx:
    for mode in ("pageable", "registered"):
        wx, wy = Image.from_array(wx0), Image.from_array(wy0)
        bufs = [a.buf, b.buf, wx.buf, wy.buf]
        if mode == "registered":
            for x in bufs:
                ctx.lib.sfgpu_host_register(x.ctypes.data, x.nbytes)
        best = 1e9
        for rep in range(6):
            wx.buf[:] = Image.from_array(wx0).buf
            wy.buf[:] = Image.from_array(wy0).buf
            t0 = time.perf_counter()
            ctx.variational(wx, wy, a, b, None)
            dt = time.perf_counter() - t0
            if rep > 0:
                best = min(best, dt)
        print("%s: %.2f ms per call (%dx%d, 5x1x30)" % (mode, best * 1e3, W, H))
        if mode == "registered":
            for x in bufs:
                ctx.lib.sfgpu_host_unregister(x.ctypes.data)
