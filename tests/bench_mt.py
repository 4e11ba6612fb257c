#!/usr/bin/env python
"""Secondary measurement (not the bench.py metric; lives under tests/ because --cpu times the oracle's reference objects): the multi-frame path on BASELINE configs 3 and 4.

    python tests/bench_mt.py [--config 3|4] [--reps N] [--cpu]

config 3: Variational_MT at 1280x1024, S=3 (5 frames), Geman-McClure eps 0.5, occlusion reasoning, 2 alternations x
          10 outer x 1 inner x 30 SOR, thresholds 1e-5
config 4: 2560x1440, 3 pyramid layers (p_scale 0.9), zero initial flow, otherwise the defaults of config 3

Times sfgpu_variational_mt through the host-buffer C ABI (uploads, pyramid, host min-cut and downloads included),
prints one JSON line per config with the executed iteration counts (the early exits of variational_mt.cpp:407,436
are data dependent), the SOR / data-term event times and the fully-fused streaming model of SURVEY 8(d)
(1468 B/px per outer iteration at S=3).  --cpu additionally times the reference objects (oracle/_ref) on ONE
outer iteration of the same window on one host core.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=0, help="3, 4 or 0 = both")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--pageable", action="store_true", help="leave the host frames pageable (default: page-locked)")
    ap.add_argument("--alter", type=int, default=2)
    ap.add_argument("--outer", type=int, default=10)
    a = ap.parse_args()
    import mt_helpers as mh
    from slowflow_b200 import Context

    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    cfgs = {
        3: dict(w=1280, h=1024, zero=False, kw=dict(niter_alter=a.alter, niter_outer=a.outer, robust_color=4, robust_color_eps=0.5)),
        4: dict(w=2560, h=1440, zero=True, kw=dict(niter_alter=a.alter, niter_outer=a.outer, robust_color=4, robust_color_eps=0.5, layers=3)),
    }
    with Context(0) as ctx:
        for cid in ([a.config] if a.config else [3, 4]):
            cf = cfgs[cid]
            ims, wx, wy = mh.window(cf["w"], cf["h"], 3, zero_flow=cf["zero"])
            p = mh.params(3, **cf["kw"])
            q = mh.clone_params(p)
            ims_n = [f.copy() for f in ims]
            ctx.normalize(ims_n, q)
            if not a.pageable:  # the caller's frames and flow planes page-locked once (sfgpu_host_register)
                for f in ims_n:
                    ctx.lib.sfgpu_host_register(f.buf.ctypes.data, f.buf.nbytes)
            best, stats, prof = None, None, None
            for rep in range(a.reps + 1):
                x, y = wx.copy(), wy.copy()
                ctx.profile_enable(True)
                ctx.profile_reset()
                ctx.synchronize()
                t0 = time.perf_counter()
                ctx.variational_mt(x, y, ims_n, q, None, None)
                dt = time.perf_counter() - t0
                pr = ctx.profile_get()
                if rep > 0 and (best is None or dt < best):
                    best, stats, prof = dt, ctx.mt_stats(), pr
            npx = cf["w"] * cf["h"]
            levels_px = npx if cid == 3 else sum(int(cf["w"] * 0.9 ** l) * int(cf["h"] * 0.9 ** l) for l in range(3))
            outer = stats.outer_iterations
            model_bytes = 1468.0 * (levels_px / max(1, stats.levels)) * outer if cid == 3 else None
            line = {
                "config": cid, "workload": "Variational_MT %dx%d S=3 %s" % (cf["w"], cf["h"], json.dumps(cf["kw"])),
                "windows_per_sec": 1.0 / best, "ms_per_window": 1e3 * best, "levels": stats.levels,
                "outer_iterations_executed": outer, "sor_calls": stats.sor_calls, "graphcut_calls": stats.graphcut_calls,
                "ms_per_outer_iteration": 1e3 * best / max(1, outer),
                "setup_ms": stats.setup_ms, "graphcut_ms": stats.graphcut_ms, "total_ms_in_lib": stats.total_ms,
                "sor_ms": prof.sor_ms, "data_ms": prof.data_ms, "kernel_launches": int(prof.kernel_launches),
                "sor_effective_gbs": 44.0 * prof.sor_pixel_sweeps / max(1e-9, prof.sor_ms * 1e-3) / 1e9,
                "data_terms_per_outer": prof.data_launches / max(1, outer),
                "fused_model_bytes": model_bytes,
                "fused_model_frac_of_peak": (model_bytes / best / 1e9 / peak) if model_bytes else None,
                "peak_gbs": peak, "host_frames": "pageable" if a.pageable else "page-locked (sfgpu_host_register)",
            }
            if not a.pageable:
                for f in ims_n:
                    ctx.lib.sfgpu_host_unregister(f.buf.ctypes.data)
            if a.cpu:
                from oracle.pyoracle import Reference, SOR_LEX
                lib = Reference().lib
                p1 = mh.params(3, **dict(cf["kw"], niter_alter=1, niter_outer=1))
                t0 = time.perf_counter()
                mh.run_cpu(lib, "sf_ref_", ims, wx, wy, p1, SOR_LEX)
                line["cpu_reference_s_per_outer_iteration_1core"] = time.perf_counter() - t0
            print(json.dumps(line))


if __name__ == "__main__":
    main()
