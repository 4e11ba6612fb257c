"""Endpoint-error metrics (utils/utils.cpp:39-108 computeEPE) used by the parity gates."""
import numpy as np


def epe(u1, v1, u2, v2, border=8):
    """mean and max endpoint difference over pixels >= `border` px from every image border."""
    d = np.sqrt((np.asarray(u1, np.float64) - u2) ** 2 + (np.asarray(v1, np.float64) - v2) ** 2)
    if border > 0:
        d = d[border:-border, border:-border]
    return float(d.mean()), float(d.max())
