#!/usr/bin/env python
"""Throughput of the sharded window driver (slowflow_b200/lib/slow_flow_gpu) on synthetic PPM frames:
    python tools/driver_bench.py [W H jets S] -- runs it with 1 and 2 host threads per GPU."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slowflow_b200 import synth  # noqa: E402

W, H, JETS, S = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (1280, 1024, 8, 3))]
steps = S - 1
n = 1 + (JETS + 2) * steps
with tempfile.TemporaryDirectory() as d:
    for k in range(n):
        f = np.clip(np.rint(synth.frame(W, H, k - steps)), 0, 255).astype(np.uint8)
        with open(os.path.join(d, "frame_%d.ppm" % k), "wb") as fh:
            fh.write(b"P6\n%d %d\n255\n" % (W, H))
            fh.write(np.ascontiguousarray(f.transpose(1, 2, 0)).tobytes())
    for tpg, extra in ((1, []), (2, []), (1, ["--no-frame-cache"])):
        out = os.path.join(d, "out%d%d" % (tpg, len(extra)))
        r = subprocess.run([os.path.join(ROOT, "slowflow_b200", "lib", "slow_flow_gpu"), "--frames", os.path.join(d, "frame_%d.ppm"),
                            "--out", out, "--start", str(steps), "--jets", str(JETS), "--S", str(S), "--threads-per-gpu", str(tpg), *extra,
                            "--occlusions", "--set", "slow_flow_occlusion_reasoning=1", "--set", "slow_flow_niter_alter=2",
                            "--set", "slow_flow_robust_color=4", "--set", "slow_flow_robust_color_eps=0.5", "--set", "16bit=0",
                            "--set", "slow_flow_smoothing=1", "--set", "slow_flow_omega_0=0", "--set", "slow_flow_omega_1=2",
                            "--set", "slow_flow_occlusion_penalty=0.1", "--set", "slow_flow_occlusion_alpha=0.1"],
                           capture_output=True, text=True)
        print("threads per GPU %d %s: %s" % (tpg, " ".join(extra), [l for l in r.stdout.splitlines() if "window loop" in l or "jets," in l or "worker" in l]), r.stderr[-300:])
