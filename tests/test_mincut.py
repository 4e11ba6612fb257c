"""Host min-cut of the occlusion labelling step (slowflow_b200/csrc/sf_gridcut.hpp, sink-rooted search forest)
against brute force and against the oracle's full Boykov-Kolmogorov restatement (oracle/sfo_gridcut.hpp).

Labels are canonical (minimal sink side of a maximum flow), so two exact algorithms must agree on EVERY pixel,
not only on the energy.  No GPU is needed: sfgpu_grid_mincut is a host-only operator twin of the labelling inside
Variational_AUX_MT::optimizeOcc (variational_aux_mt.cpp:851-881)."""
import ctypes as C
import itertools

import numpy as np
import pytest

from slowflow_b200.api import load_library

pytestmark = pytest.mark.usefixtures("built")
FP, IP = C.POINTER(C.c_float), C.POINTER(C.c_int)


def product_cut(w, h, d0, d1, alpha, int_terms=0):
    lib = load_library()
    lab = np.full(w * h, -7, np.int32)
    rc = lib.sfgpu_grid_mincut(w, h, d0.ctypes.data_as(FP), d1.ctypes.data_as(FP), C.c_float(alpha), int_terms,
                               lab.ctypes.data_as(IP))
    assert rc == 0
    return lab


def oracle_cut(oracle, w, h, d0, d1, alpha, int_terms=0):
    L = oracle.lib
    L.sfo_mincut.argtypes = [C.c_int, C.c_int, FP, FP, C.c_float, C.c_int, IP]
    lab = np.full(w * h, -7, np.int32)
    assert L.sfo_mincut(w, h, d0.ctypes.data_as(FP), d1.ctypes.data_as(FP), C.c_float(alpha), int_terms,
                        lab.ctypes.data_as(IP)) == 0
    return lab


def test_mincut_brute_force():
    r = np.random.RandomState(3)
    w, h = 4, 3
    q = lambda v: np.round(np.asarray(v, np.float64) * 2 ** 24) / 2 ** 24

    def energy(lab, d0, d1, a):
        e = sum(d1[p] if lab[p] else d0[p] for p in range(w * h))
        for y in range(h):
            for x in range(w):
                p = y * w + x
                e += a * (int(x + 1 < w and lab[p] != lab[p + 1]) + int(y + 1 < h and lab[p] != lab[p + w]))
        return e

    for _ in range(80):
        d0, d1 = r.rand(w * h).astype(np.float32), r.rand(w * h).astype(np.float32)
        a = np.float32(r.rand() * 0.5)
        lab = product_cut(w, h, d0, d1, float(a))
        assert set(np.unique(lab)) <= {0, 1}
        best = min(energy(l, q(d0), q(d1), float(q(a))) for l in itertools.product([0, 1], repeat=w * h))
        assert abs(energy(lab, q(d0), q(d1), float(q(a))) - best) < 1e-9
    # integer EnergyTermType (stock gco): every cost < 1 truncates to 0 -> nothing leaves label 0
    assert not product_cut(w, h, d0, d1, float(a), int_terms=1).any()


@pytest.mark.parametrize("w,h,seed,kind", [
    (37, 23, 1, "uniform"), (64, 48, 2, "uniform"), (131, 77, 3, "occlusion"), (200, 150, 4, "occlusion"),
    (97, 61, 5, "ties"), (1, 40, 6, "uniform"), (40, 1, 7, "uniform"), (160, 120, 8, "blobs"),
])
def test_mincut_matches_oracle_labels(oracle, w, h, seed, kind):
    """Every pixel's label equals the oracle's (canonical cut), on the cost statistics the path produces:
    'occlusion' = tiny label-0 costs, label-1 costs = penalty 0.1 + tiny, a few strongly occluded blobs."""
    r = np.random.RandomState(seed)
    n = w * h
    if kind == "uniform":
        d0, d1, alpha = r.rand(n), r.rand(n), 0.3 * r.rand()
    elif kind == "ties":  # many equal costs: exercises the canonical (minimal sink side) choice
        d0, d1, alpha = r.randint(0, 3, n) * 0.25, r.randint(0, 3, n) * 0.25, 0.25
    else:
        d0 = 0.002 * r.rand(n)
        d1 = 0.1 + 0.002 * r.rand(n)
        yy, xx = np.mgrid[0:h, 0:w]
        for _ in range(6 if kind == "occlusion" else 25):
            cx, cy, rad = r.randint(0, w), r.randint(0, h), r.randint(2, max(3, min(w, h) // 5))
            m = ((xx - cx) ** 2 + (yy - cy) ** 2 <= rad * rad).ravel()
            d0[m] += r.rand() * (0.6 if kind == "occlusion" else 0.25)
        alpha = 0.1
    d0, d1 = d0.astype(np.float32), d1.astype(np.float32)
    a = product_cut(w, h, d0, d1, float(alpha))
    b = oracle_cut(oracle, w, h, d0, d1, float(alpha))
    assert np.array_equal(a, b), "labels differ at %d of %d pixels" % (int((a != b).sum()), n)
    if kind in ("occlusion", "blobs"):
        assert 0 < a.sum() < n  # both labels occur


def test_reused_cut_object_equals_fresh_ones(tmp_path):
    """tests/cpp_mincut_reuse.cpp: 300 solves on one reused object vs fresh objects (sizes and capacities change)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "reuse")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(root, "tests", "cpp_mincut_reuse.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout + r.stderr


# ------------------------------------------------------------------ the DEVICE solver (sf_mincut.cu) -- needs a GPU
def device_cut(ctx, w, h, d0, d1, alpha, int_terms=0):
    lab = np.full(w * h, -7, np.int32)
    stats = (C.c_int * 2)()
    rc = ctx.lib.sfgpu_grid_mincut_dev(ctx.h, w, h, d0.ctypes.data_as(FP), d1.ctypes.data_as(FP), C.c_float(alpha), int_terms,
                                       lab.ctypes.data_as(IP), stats)
    assert rc == 0, ctx.lib.sfgpu_last_error().decode()
    return lab, (stats[0], stats[1])


def _costs(w, h, seed, kind):
    r = np.random.RandomState(seed)
    n = w * h
    if kind == "uniform":
        d0, d1, alpha = r.rand(n), r.rand(n), 0.3 * r.rand()
    elif kind == "ties":
        d0, d1, alpha = r.randint(0, 3, n) * 0.25, r.randint(0, 3, n) * 0.25, 0.25
    elif kind == "no_pairwise":
        d0, d1, alpha = r.rand(n), r.rand(n), 0.0
    elif kind == "all_sink":  # every pixel prefers label 1: the whole image is excess, nothing can be absorbed
        d0, d1, alpha = 0.5 + r.rand(n), 0.1 * r.rand(n), 0.2
    elif kind == "wide_band":  # a 60-px wide occluded band: long residual paths (many relaxation sweeps)
        d0, d1, alpha = 0.002 * r.rand(n), 0.1 + 0.002 * r.rand(n), 0.1
        xx = np.mgrid[0:h, 0:w][1].ravel()
        d0[(xx > w // 2 - 30) & (xx < w // 2 + 30)] += 0.103 + 0.01 * r.rand(int(((xx > w // 2 - 30) & (xx < w // 2 + 30)).sum()))
    else:
        d0 = 0.002 * r.rand(n)
        d1 = 0.1 + 0.002 * r.rand(n)
        yy, xx = np.mgrid[0:h, 0:w]
        for _ in range(6 if kind == "occlusion" else 25):
            cx, cy, rad = r.randint(0, w), r.randint(0, h), r.randint(2, max(3, min(w, h) // 5))
            m = ((xx - cx) ** 2 + (yy - cy) ** 2 <= rad * rad).ravel()
            d0[m] += r.rand() * (0.6 if kind == "occlusion" else 0.25)
        alpha = 0.1
    return d0.astype(np.float32), d1.astype(np.float32), float(alpha)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,seed,kind", [
    (37, 23, 1, "uniform"), (64, 48, 2, "uniform"), (131, 77, 3, "occlusion"), (200, 150, 4, "occlusion"),
    (97, 61, 5, "ties"), (1, 40, 6, "uniform"), (40, 1, 7, "uniform"), (160, 120, 8, "blobs"), (90, 70, 9, "no_pairwise"),
    (120, 80, 10, "all_sink"), (400, 96, 11, "wide_band"), (1280, 1024, 12, "occlusion"), (1280, 1024, 13, "blobs"),
    (2560, 1440, 14, "blobs"),
])
def test_device_mincut_labels_equal_host_solver(ctx, w, h, seed, kind):
    """The cooperative-kernel solver the multi-frame path runs (no D2H of capacities, no host cut) labels every pixel
    like the host SinkForestCut -- which test_mincut_matches_oracle_labels pins to the oracle's Boykov-Kolmogorov."""
    d0, d1, alpha = _costs(w, h, seed, kind)
    host = product_cut(w, h, d0, d1, alpha)
    dev, (phases, passes) = device_cut(ctx, w, h, d0, d1, alpha)
    print("%dx%d %s: %d phases, %d grid-wide passes, %d pixels labelled 1" % (w, h, kind, phases, passes, int(dev.sum())))
    assert np.array_equal(host, dev), "labels differ at %d of %d pixels" % (int((host != dev).sum()), w * h)


@pytest.mark.gpu
def test_device_mincut_brute_force_and_int_terms(ctx):
    r = np.random.RandomState(5)
    w, h = 4, 3
    for _ in range(40):
        d0, d1 = r.rand(w * h).astype(np.float32), r.rand(w * h).astype(np.float32)
        a = float(np.float32(r.rand() * 0.5))
        assert np.array_equal(device_cut(ctx, w, h, d0, d1, a)[0], product_cut(w, h, d0, d1, a))
    assert not device_cut(ctx, w, h, d0, d1, a, int_terms=1)[0].any()


@pytest.mark.gpu
def test_mt_device_cut_equals_host_cut_path(monkeypatch):
    """The whole multi-frame solve with the occlusion labelling on the device (default) and on the host
    (SLOWFLOW_GPU_HOST_MINCUT=1): identical labels, identical flow."""
    import mt_helpers as mh
    from slowflow_b200 import Context
    ims, wx, wy = mh.window(320, 200, 3)
    p = mh.params(3, niter_alter=3, niter_outer=3, robust_color=4, robust_color_eps=0.5)
    with Context(0) as dev_ctx:
        a = mh.run_gpu(dev_ctx, ims, wx, wy, p)
    monkeypatch.setenv("SLOWFLOW_GPU_HOST_MINCUT", "1")
    with Context(0) as host_ctx:
        b = mh.run_gpu(host_ctx, ims, wx, wy, p)
    assert a["stats"].graphcut_calls == 2 and b["stats"].graphcut_calls == 2
    assert np.array_equal(a["occ"].array, b["occ"].array)
    assert np.array_equal(a["wx"].array, b["wx"].array) and np.array_equal(a["wy"].array, b["wy"].array)
    assert (a["occ"].array == -1).any()
