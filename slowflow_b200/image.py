"""ctypes mirrors of the reference containers (epic_flow_extended/image.h:17-34) backed by numpy.

``Image`` / ``ColorImage`` own a 64-byte aligned float32 buffer with ``stride = ceil4(width)``
(image.c:25) and expose both a numpy view and the C struct that crosses the ABI.
"""
import ctypes as C

import numpy as np


class image_t(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("stride", C.c_int), ("data", C.POINTER(C.c_float))]


class color_image_t(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("stride", C.c_int),
                ("c1", C.POINTER(C.c_float)), ("c2", C.POINTER(C.c_float)), ("c3", C.POINTER(C.c_float))]


def ceil4(w):
    return ((int(w) + 3) // 4) * 4


def _aligned(n_floats, align=64):
    raw = np.zeros(n_floats * 4 + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n_floats * 4].view(np.float32)


class Image:
    """One float plane (reference ``image_t``)."""

    def __init__(self, width, height, buffer=None):
        self.width, self.height, self.stride = int(width), int(height), ceil4(width)
        n = self.stride * self.height
        self.buf = _aligned(n) if buffer is None else buffer
        assert self.buf.dtype == np.float32 and self.buf.size == n and self.buf.ctypes.data % 16 == 0
        self.c = image_t(self.width, self.height, self.stride, self.buf.ctypes.data_as(C.POINTER(C.c_float)))

    @classmethod
    def from_array(cls, a):
        a = np.asarray(a, dtype=np.float32)
        im = cls(a.shape[1], a.shape[0])
        im.array[:] = a
        return im

    @property
    def full(self):  # (height, stride) including the padding columns
        return self.buf.reshape(self.height, self.stride)

    @property
    def array(self):  # (height, width) valid pixels
        return self.full[:, :self.width]

    def copy(self):
        o = Image(self.width, self.height)
        o.buf[:] = self.buf
        return o

    def ptr(self):
        return C.byref(self.c)


class ColorImage:
    """Three planar float channels (reference ``color_image_t``; c2 = c1 + stride*height)."""

    def __init__(self, width, height, buffer=None):
        self.width, self.height, self.stride = int(width), int(height), ceil4(width)
        n = 3 * self.stride * self.height
        self.buf = _aligned(n) if buffer is None else buffer
        assert self.buf.dtype == np.float32 and self.buf.size == n and self.buf.ctypes.data % 16 == 0
        base = self.buf.ctypes.data
        plane = self.stride * self.height * 4
        P = C.POINTER(C.c_float)
        self.c = color_image_t(self.width, self.height, self.stride, C.cast(base, P), C.cast(base + plane, P),
                               C.cast(base + 2 * plane, P))

    @classmethod
    def from_array(cls, a):
        """a: (3, H, W) float array."""
        a = np.asarray(a, dtype=np.float32)
        im = cls(a.shape[2], a.shape[1])
        im.array[:] = a
        return im

    @property
    def full(self):
        return self.buf.reshape(3, self.height, self.stride)

    @property
    def array(self):
        return self.full[:, :, :self.width]

    def copy(self):
        o = ColorImage(self.width, self.height)
        o.buf[:] = self.buf
        return o

    def ptr(self):
        return C.byref(self.c)
