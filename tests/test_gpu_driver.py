"""The sharded window driver (tools/slow_flow_gpu.cpp -> slowflow_b200/lib/slow_flow_gpu): the reference's jet loop
(slow_flow.cpp:706-1030, deep_matching = 0) on one host thread per GPU, written against the C++ shim.  Its .flo / .pbm
outputs must equal what the same windows give through the Python mirror of the ABI (same library: bit-identical)."""
import os
import subprocess

import numpy as np
import pytest

import mt_helpers as mh
from slowflow_b200 import ColorImage, Image, Variational_MT, mt_params_default, synth
from slowflow_b200.api import load_library, read_flo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "slowflow_b200", "lib", "slow_flow_gpu")

# ParameterList key -> sf_mt_params_t field (include/variational_mt_gpu.hpp)
KEYS = [("slow_flow_layers", "layers"), ("slow_flow_p_scale", "p_scale"), ("slow_flow_alpha", "alpha"),
        ("slow_flow_gamma", "gamma"), ("slow_flow_delta", "delta"), ("slow_flow_dataterm", "dataterm"),
        ("slow_flow_smoothing", "smoothing"), ("slow_flow_robust_color", "robust_color"),
        ("slow_flow_robust_color_eps", "robust_color_eps"), ("slow_flow_robust_color_truncation", "robust_color_truncation"),
        ("slow_flow_robust_reg", "robust_reg"), ("slow_flow_robust_reg_eps", "robust_reg_eps"),
        ("slow_flow_robust_reg_truncation", "robust_reg_truncation"), ("slow_flow_niter_alter", "niter_alter"),
        ("slow_flow_niter_outer", "niter_outer"), ("slow_flow_niter_inner", "niter_inner"),
        ("slow_flow_niter_solver", "niter_solver"), ("slow_flow_niter_graphc", "niter_graphc"),
        ("slow_flow_thres_outer", "thres_outer"), ("slow_flow_thres_inner", "thres_inner"),
        ("slow_flow_sor_omega", "sor_omega"), ("slow_flow_occlusion_reasoning", "occlusion_reasoning"),
        ("slow_flow_occlusion_penalty", "occlusion_penalty"), ("slow_flow_occlusion_alpha", "occlusion_alpha"),
        ("16bit", "hbit")]


def write_ppm(path, rgb_u8):
    h, w, _ = rgb_u8.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(rgb_u8.tobytes())


@pytest.mark.parametrize("S,jets,gpus,hbit", [(2, 3, 1, 0), (3, 2, 1, 0), (2, 4, 2, 0), (2, 2, 1, 1)])
def test_driver_outputs_equal_api_results(ctx, tmp_path, S, jets, gpus, hbit):
    lib = load_library()
    if gpus > lib.sfgpu_device_count():
        pytest.skip("needs %d GPUs" % gpus)
    w, h, start, steps = 150, 98, 5, S - 1
    nframes = 1 + (jets + 2) * steps
    seqdir, out = tmp_path / "seq", tmp_path / "out"
    seqdir.mkdir()
    frames_u8 = []
    for k in range(nframes):
        f = np.clip(np.rint(synth.frame(w, h, k - steps)), 0, 255).astype(np.uint8)  # (3, H, W)
        frames_u8.append(f)
        write_ppm(seqdir / ("frame_%d.ppm" % (start - steps + k)), np.ascontiguousarray(f.transpose(1, 2, 0)))
    p = mt_params_default()
    p.S, p.hbit, p.niter_alter, p.niter_outer = S, hbit, 2, 3  # hbit = 1: the shipped cfg default (16bit)
    for a in range(S - 1):
        p.rho[a], p.omega[a] = 1.0, 1.0 + a
    args = [DRIVER, "--frames", str(seqdir / "frame_%d.ppm"), "--out", str(out), "--start", str(start), "--jets", str(jets),
            "--S", str(S), "--gpus", str(gpus), "--occlusions"]
    for key, field in KEYS:
        args += ["--set", "%s=%r" % (key, getattr(p, field))]
    for a in range(S - 1):
        args += ["--set", "slow_flow_rho_%d=%r" % (a, p.rho[a]), "--set", "slow_flow_omega_%d=%r" % (a, p.omega[a])]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "%d jets" % jets in r.stdout

    # the same windows through the Python mirror
    ims = [ColorImage.from_array(f.astype(np.float32)) for f in frames_u8]
    q = mh.clone_params(p)
    ctx.normalize(ims, q)
    back = ims[::-1]
    for j in range(jets):
        f = j * steps
        for name, window, idx in (("frame_%d.flo", ims[f:f + 2 * steps + 1], start + f),
                                  ("frame_%d_back.flo", back[nframes - 1 - f - 3 * steps:][:2 * steps + 1], start + f + steps)):
            solver = Variational_MT(ctx)
            ones = ColorImage(w, h)
            ones.buf[:] = 1.0
            solver.setChannelWeights(ones)
            wx, wy = Image(w, h), Image(w, h)
            wx.buf[:] = 0
            wy.buf[:] = 0
            solver.variational(wx, wy, window, q)
            gx, gy = read_flo(out / (name % idx))
            assert np.array_equal(gx.array, wx.array * steps) and np.array_equal(gy.array, wy.array * steps), (j, name)
            if name.endswith("%d.flo"):
                pbm = (out / "occlusion" / ("frame_%d.pbm" % idx)).read_bytes()
                bits = np.packbits(solver.getOcclusions().array == -1, axis=1)
                assert pbm == b"P4\n%d %d\n" % (w, h) + bits.tobytes()
    cfg = (out / "config.cfg").read_text()
    assert "slow_flow_S\t%d" % S in cfg and "Jets\t\t%d" % jets in cfg


def test_driver_resume_skips_finished_windows(tmp_path):
    """-resume (slow_flow.cpp:178, 794, 958): a window whose .flo exists is not recomputed; a missing one is."""
    w, h, start, S, jets = 96, 64, 3, 2, 2
    steps = S - 1
    seqdir, out = tmp_path / "seq", tmp_path / "out"
    seqdir.mkdir()
    for k in range(1 + (jets + 2) * steps):
        f = np.clip(np.rint(synth.frame(w, h, k - steps)), 0, 255).astype(np.uint8)
        write_ppm(seqdir / ("frame_%d.ppm" % (start - steps + k)), np.ascontiguousarray(f.transpose(1, 2, 0)))
    args = [DRIVER, "--frames", str(seqdir / "frame_%d.ppm"), "--out", str(out), "--start", str(start), "--jets", str(jets),
            "--S", str(S), "--gpus", "1", "--set", "slow_flow_niter_alter=1", "--set", "slow_flow_niter_outer=2"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    first = {p.name: p.read_bytes() for p in out.glob("*.flo")}
    assert len(first) == 2 * jets
    victim = out / ("frame_%d_back.flo" % (start + steps))
    victim.unlink()
    marker = out / ("frame_%d.flo" % start)
    marker.write_bytes(b"kept")  # a finished file is never touched, whatever it holds
    r = subprocess.run(args + ["--resume"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "3 finished flow files skipped" in r.stdout
    assert marker.read_bytes() == b"kept"
    assert victim.read_bytes() == first[victim.name]
