"""Multi-frame golden vectors from the REFERENCE's unmodified driver (oracle/_ref: normalize +
Variational_MT::variational).  Called by make_golden.py."""
import os

import numpy as np

import mt_helpers as mh

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "mt_S3_96x64_modl1_occ": dict(size=(96, 64), S=3, kw=dict(niter_alter=2, niter_outer=3)),
    "mt_S3_77x61_gm_pyr2": dict(size=(77, 61), S=3, zero=True, kw=dict(layers=2, niter_alter=1, niter_outer=2, robust_color=4,
                                                                      robust_color_eps=0.5)),
    "mt_S2_64x48_lorentzian_noocc": dict(size=(64, 48), S=2, kw=dict(niter_alter=1, niter_outer=3, robust_color=2,
                                                                     robust_color_eps=0.5, occlusion_reasoning=0, niter_inner=2)),
}


def main(ref):
    for name, c in CASES.items():
        ims, wx, wy = mh.window(c["size"][0], c["size"][1], c["S"], zero_flow=c.get("zero", False))
        p = mh.params(c["S"], **c["kw"])
        out = {}
        for mode, tag in ((0, "lex"), (1, "rb")):
            r = mh.run_cpu(ref.lib, "sf_ref_", ims, wx, wy, p, mode)
            out["wx_" + tag], out["wy_" + tag], out["occ_" + tag] = r["wx"].array.copy(), r["wy"].array.copy(), r["occ"].array.copy()
            out["avg_" + tag] = np.array(r["avg"], np.float32)
        out["norm_avg"] = np.array([r["params"].img_norm_avg[k] for k in range(3)], np.float32)
        out["norm_std"] = np.array([r["params"].img_norm_std[k] for k in range(3)], np.float32)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print("wrote", name)
