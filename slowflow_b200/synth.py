"""Synthetic inputs of SURVEY.md section 8(d): analytic textures, a ground-truth flow with one motion
discontinuity, backward-warped frames at any temporal offset and a noisy initial flow.

Everything is a closed-form function of (x, y, seed): frames are exact at sub-pixel displacements,
reproducible on any host, and need no data files (SED edges, DeepMatching and the teaser data are
unavailable offline).
"""
import numpy as np

_WAVELENGTHS = (7.0, 11.0, 19.0, 37.0, 71.0, 131.0, 3.5, 5.0)
_AMPLITUDES = (20.0, 20.0, 25.0, 25.0, 30.0, 30.0, 8.0, 8.0)


class _PCG32:
    """Minimal PCG32 (XSH-RR) so the texture parameters do not depend on numpy's generator versions."""

    def __init__(self, seed, seq=54):
        self.state, self.inc = 0, ((seq << 1) | 1) & 0xFFFFFFFFFFFFFFFF
        self.next()
        self.state = (self.state + seed) & 0xFFFFFFFFFFFFFFFF
        self.next()

    def next(self):
        old = self.state
        self.state = (old * 6364136223846793005 + self.inc) & 0xFFFFFFFFFFFFFFFF
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF

    def uniform(self):
        return self.next() / 4294967296.0


def texture_params(seed=20170721):
    """theta_k (shared by the channels) and phi_{k,c}."""
    rng = _PCG32(seed)
    theta = [rng.uniform() * np.pi for _ in _WAVELENGTHS]
    phi = []
    for c in range(3):
        rc = _PCG32(seed + c + 1)
        phi.append([rc.uniform() * 2 * np.pi for _ in _WAVELENGTHS])
    return np.array(theta), np.array(phi)


def texture(x, y, seed=20170721):
    """T_c(x, y) for float coordinate arrays -> (3, ...) in [0, 255]."""
    theta, phi = texture_params(seed)
    out = np.full((3,) + x.shape, 127.5, dtype=np.float64)
    for k, (lam, amp) in enumerate(zip(_WAVELENGTHS, _AMPLITUDES)):
        f = 2 * np.pi / lam
        arg = f * (x * np.cos(theta[k]) + y * np.sin(theta[k]))
        for c in range(3):
            out[c] += amp * 0.68 * np.sin(arg + phi[c, k])
    return out


def gt_flow(width, height):
    """Ground-truth flow per frame step: smooth + one vertical 3 px motion discontinuity at x = W/2."""
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    u = 1.0 + 2.0 * np.sin(2 * np.pi * y / height * 1.5) + (x > width / 2) * 3.0
    v = 1.5 * np.cos(2 * np.pi * x / width * 2)
    return u, v


def frame(width, height, t, seed=20170721):
    """I_t(x,y) = T(x - t*u, y - t*v) -> float32 (3, H, W)."""
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    u, v = gt_flow(width, height)
    return texture(x - t * u, y - t * v, seed).astype(np.float32)


def _hash_uniform(n, seed):
    """Counter-based uniform [0,1) (splitmix64 finaliser), identical on every platform."""
    base = (int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) + np.uint64(base)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def initial_flow(width, height, noise=0.25, seed=777, zero=False):
    """GT + U(-noise, noise) per pixel (configs 1-3) or zeros (config 4)."""
    if zero:
        return np.zeros((height, width), np.float32), np.zeros((height, width), np.float32)
    u, v = gt_flow(width, height)
    n = width * height
    nu = (_hash_uniform(n, seed) * 2 - 1).reshape(height, width) * noise
    nv = (_hash_uniform(n, seed + 1) * 2 - 1).reshape(height, width) * noise
    return (u + nu).astype(np.float32), (v + nv).astype(np.float32)


def two_frame_case(width, height, seed=20170721, noise=0.25):
    """(im1, im2, wx0, wy0) as float32 arrays for the two-frame configs."""
    im1, im2 = frame(width, height, 0, seed), frame(width, height, 1, seed)
    wx, wy = initial_flow(width, height, noise)
    return im1, im2, wx, wy


def window_case(width, height, S=3, seed=20170721, noise=0.25, zero_flow=False):
    """F = 2(S-1)+1 frames at t = -ref..ref plus the initial flow per frame step."""
    ref = S - 1
    frames = [frame(width, height, t, seed) for t in range(-ref, ref + 1)]
    wx, wy = initial_flow(width, height, noise, zero=zero_flow)
    return frames, wx, wy


def epic_case(width, height, n_matches=2000, seed=4242, outliers=0.05):
    """Synthetic inputs of the EPIC interpolation (epic.cpp:147): the first frame (used for the saliency filter), sparse
    matches x1 y1 x2 y2 scattered over the image that follow the ground-truth flow up to +-0.3 px (plus a few gross
    outliers for the consistency filter), and an edge-cost map in [0, 1] that is high along the motion discontinuity and
    along the texture's strongest gradients (the reference takes SED edges here, which are not available offline)."""
    rng = np.random.RandomState(seed)
    im = frame(width, height, 0)
    u, v = gt_flow(width, height)
    xs = rng.randint(2, width - 2, n_matches)
    ys = rng.randint(2, height - 2, n_matches)
    m = np.zeros((n_matches, 4), np.float32)
    m[:, 0], m[:, 1] = xs, ys
    m[:, 2] = xs + u[ys, xs] + rng.uniform(-0.3, 0.3, n_matches)
    m[:, 3] = ys + v[ys, xs] + rng.uniform(-0.3, 0.3, n_matches)
    bad = rng.rand(n_matches) < outliers
    m[bad, 2] += rng.uniform(-25, 25, int(bad.sum()))
    m[bad, 3] += rng.uniform(-25, 25, int(bad.sum()))
    lum = im.mean(axis=0)
    gy, gx = np.gradient(lum)
    mag = np.sqrt(gx * gx + gy * gy)
    edges = 0.25 * mag / (mag.max() + 1e-9)
    x = np.arange(width)[None, :]
    edges = edges + 0.9 * np.exp(-0.5 * ((x - width / 2.0) / 1.5) ** 2)
    return im, m, np.clip(edges, 0.0, 1.0).astype(np.float32)
