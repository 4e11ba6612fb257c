"""Debug helper: run the SOR operator twin on small cases and print the CUDA error, if any."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers
from slowflow_b200 import Context, Image
from oracle.pyoracle import Oracle, SOR_REDBLACK

orc = Oracle()
for (w, h, variant, fuse) in [(64, 64, 1, 1), (64, 64, 0, 1), (20, 17, 0, 1), (131, 77, 0, 3), (300, 200, 0, 5)]:
    try:
        ctx = Context(0)
        im1, im2, wx, wy = helpers.pair(w, h)
        s = helpers.oracle_system(orc, im1, im2, wx, wy)
        du0, dv0 = helpers.rng_plane(w, h, 11, -0.2, 0.2), helpers.rng_plane(w, h, 12, -0.2, 0.2)
        Ar = [a.copy() for a in s["A"]]
        rdu, rdv = du0.copy(), dv0.copy()
        orc.lib.sfo_sor_coupled(rdu.ptr(), rdv.ptr(), *[a.ptr() for a in Ar], s["sh"].ptr(), s["sv"].ptr(), 30, 1.9, SOR_REDBLACK)
        ctx.set_sor_variant(variant); ctx.set_sor_fuse(fuse)
        Ag = [a.copy() for a in s["A"]]
        gdu, gdv = du0.copy(), dv0.copy()
        ctx.sor_coupled(gdu, gdv, *Ag, s["sh"], s["sv"], 30, 1.9)
        d = np.abs(gdu.array - rdu.array)
        print("OK %dx%d variant %d fuse %d: max diff du %.3e dv %.3e; worst at %s" % (
            w, h, variant, fuse, d.max(), np.abs(gdv.array - rdv.array).max(), np.unravel_index(d.argmax(), d.shape)), flush=True)
        ctx.close()
    except Exception as e:
        print("FAIL %dx%d variant %d fuse %d: %s" % (w, h, variant, fuse, e), flush=True)
        break
