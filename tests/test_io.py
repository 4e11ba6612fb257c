"""On-disk outputs of the flow stage: .flo files (epic_flow_extended/io.c:50-96) and the occlusion
.pbm of slow_flow.cpp:893-905.  Host-only entry points of the library; no GPU needed."""
import ctypes as C

import numpy as np
import pytest

from slowflow_b200 import Image
from slowflow_b200.api import read_flo, write_flo, write_occlusion_pbm
from slowflow_b200.image import image_t

pytestmark = pytest.mark.usefixtures("built")


@pytest.mark.parametrize("w,h", [(64, 48), (61, 45), (1, 7), (130, 3)])
def test_flo_bytes_match_reference_writer_and_round_trip(oracle, tmp_path, w, h):
    r = np.random.RandomState(w * 100 + h)
    wx, wy = Image.from_array(r.randn(h, w).astype(np.float32) * 5), Image.from_array(r.randn(h, w).astype(np.float32) * 5)
    wx.buf.reshape(h, -1)[:, w:] = 777.0  # stride padding must not leak into the file
    a, b = tmp_path / "a.flo", tmp_path / "b.flo"
    write_flo(a, wx, wy)
    L = oracle.lib
    L.sfo_write_flo.argtypes = [C.c_char_p, C.POINTER(image_t), C.POINTER(image_t)]
    assert L.sfo_write_flo(str(b).encode(), wx.ptr(), wy.ptr()) == 0
    raw = a.read_bytes()
    assert raw == b.read_bytes()
    assert len(raw) == 12 + 8 * w * h and np.frombuffer(raw[:4], np.float32)[0] == np.float32(202021.25)
    rx, ry = read_flo(a)
    assert np.array_equal(rx.array, wx.array) and np.array_equal(ry.array, wy.array)


@pytest.mark.parametrize("w,h", [(64, 48), (61, 45), (1, 7)])
def test_flo_against_the_reference_io_functions(reference, tmp_path, w, h):
    """writeFlowFile / readFlowFile of the reference's own io.c (:50-96), compiled into oracle/_ref: our writer's bytes
    equal the reference writer's, and the reference reader reads our file back."""
    r = np.random.RandomState(w * 31 + h)
    wx, wy = Image.from_array(r.randn(h, w).astype(np.float32) * 5), Image.from_array(r.randn(h, w).astype(np.float32) * 5)
    wx.buf.reshape(h, -1)[:, w:] = 777.0
    a, b = tmp_path / "ours.flo", tmp_path / "ref.flo"
    write_flo(a, wx, wy)
    L = reference.lib
    IP = C.POINTER(image_t)
    L.writeFlowFile.argtypes = [C.c_char_p, IP, IP]
    L.writeFlowFile.restype = None
    L.writeFlowFile(str(b).encode(), wx.ptr(), wy.ptr())
    assert a.read_bytes() == b.read_bytes()
    L.readFlowFile.argtypes = [C.c_char_p]
    L.readFlowFile.restype = C.POINTER(IP)
    flow = L.readFlowFile(str(a).encode())  # (leaks two small images: the reference's own ownership rule)
    for k, src in enumerate((wx, wy)):
        im = flow[k].contents
        assert (im.width, im.height) == (w, h)
        got = np.ctypeslib.as_array(im.data, shape=(h, im.stride))[:, :w]
        assert np.array_equal(got, src.array)
    ox, oy = read_flo(b)  # and our reader reads the reference writer's file
    assert np.array_equal(ox.array, wx.array) and np.array_equal(oy.array, wy.array)


def test_flo_reader_rejects_garbage(tmp_path):
    p = tmp_path / "bad.flo"
    p.write_bytes(b"PIEH" + b"\0" * 20)
    with pytest.raises(RuntimeError):
        read_flo(p)


@pytest.mark.parametrize("w,h", [(64, 48), (61, 45), (9, 2), (8, 5)])
def test_occlusion_pbm_matches_opencv(oracle, tmp_path, w, h):
    r = np.random.RandomState(w + h)
    occ = Image.from_array(r.randint(-1, 2, size=(h, w)).astype(np.float32))
    p = tmp_path / "occ.pbm"
    write_occlusion_pbm(p, occ)
    raw = p.read_bytes()
    # the oracle's 8-bit image (0 / 128 / 255), packed the PBM way: bit set <=> value 0
    L = oracle.lib
    L.sfo_occlusion_to_u8.argtypes = [C.POINTER(image_t), C.c_void_p]
    u8 = np.zeros((h, w), np.uint8)
    assert L.sfo_occlusion_to_u8(occ.ptr(), u8.ctypes.data) == 0
    assert set(np.unique(u8)) <= {0, 128, 255}
    bits = np.packbits(u8 == 0, axis=1)
    assert raw == b"P4\n%d %d\n" % (w, h) + bits.tobytes()
    cv2 = pytest.importorskip("cv2")
    q = tmp_path / "cv.pbm"
    assert cv2.imwrite(str(q), u8, [cv2.IMWRITE_PXM_BINARY, 1])
    assert raw == q.read_bytes()
    back = cv2.imread(str(p), cv2.IMREAD_UNCHANGED)  # what dense_tracking.cpp reads: 0 where occluded in the past
    assert np.array_equal(back == 0, occ.array == -1)
