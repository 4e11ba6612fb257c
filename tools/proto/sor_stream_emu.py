#!/usr/bin/env python
"""Protocol emulation of the streaming red-black SOR kernel (sf_sor_stream.cu): 9 warps, each owning two row slots
(stream rows 2w, 2w+1 modulo 18), one row relaxed per turn, level k of row j at step j + 1 + 2k, published boundary
data double-buffered by version, neighbour-only step counters.  The scheduler below picks ANY warp whose wait condition
holds, so protocol races show up as a mismatch against plain red-black half sweeps.  numpy float64, rows as vectors."""
import sys
import numpy as np

NW, WIN = 9, 18


def reference(co, du, dv, nlev, omega):
    n, wd = du.shape
    du, dv = du.copy(), dv.copy()
    yy, xx = np.mgrid[0:n, 0:wd]
    for k in range(nlev):
        m = ((xx + yy) & 1) == (k & 1)
        B = []
        for d in (du, dv):
            l = np.zeros_like(d); l[:, 1:] = d[:, :-1]
            r = np.zeros_like(d); r[:, :-1] = d[:, 1:]
            t = np.zeros_like(d); t[1:] = d[:-1]
            b = np.zeros_like(d); b[:-1] = d[1:]
            B.append((l, r, t, b))
        psl = np.zeros_like(du); psl[:, 1:] = co['ph'][:, :-1]
        pst = np.zeros_like(du); pst[1:] = co['pv'][:-1]
        B1 = psl * B[0][0] + co['ph'] * B[0][1] + pst * B[0][2] + co['pv'] * B[0][3] + co['b1']
        B2 = psl * B[1][0] + co['ph'] * B[1][1] + pst * B[1][2] + co['pv'] * B[1][3] + co['b2']
        nu = du + omega * (co['a11'] * B1 + co['a12'] * B2 - du)
        nv = dv + omega * (co['a12'] * B1 + co['a22'] * B2 - dv)
        du = np.where(m, nu, du)
        dv = np.where(m, nv, dv)
    return du, dv


class Warp:
    def __init__(self, w):
        self.w = w
        self.s = -1            # next step to execute
        self.row = [None, None]  # stream row held by each slot
        self.du = [None, None]
        self.dv = [None, None]


def run(n, wd, nlev, seed, omega=1.9):
    rng = np.random.RandomState(seed)
    co = {k: rng.rand(n, wd) * 0.2 for k in ('a11', 'a12', 'a22', 'b1', 'b2', 'ph', 'pv')}
    co['ph'][:, -1] = 0
    du0, dv0 = rng.randn(n, wd), rng.randn(n, wd)
    ref_u, ref_v = reference(co, du0, dv0, nlev, omega)
    out_u, out_v = np.full((n, wd), np.nan), np.full((n, wd), np.nan)
    # published areas: [residue][buffer][parity] -> (du, dv) of that column parity
    pub = [[[(np.zeros((wd + 1 - p) // 2), np.zeros((wd + 1 - p) // 2)) for p in range(2)] for _ in range(2)] for _ in range(WIN)]
    cnt = [-2] * NW
    warps = [Warp(w) for w in range(NW)]
    last = n + 16

    def ready(W):
        s = W.s
        b = (s + 1) & 1
        if b == 1:
            return cnt[(W.w + 1) % NW] >= s - 1
        return cnt[(W.w - 1) % NW] >= s - 3

    def turn(W):
        s, w = W.s, W.w
        b = (s + 1) & 1
        phi = ((s - 1 - 2 * w - b) // 2) % 9
        if phi < 8:
            j = s - 1 - 2 * phi
            k = phi
            if 0 <= j < n and k < nlev:
                assert W.row[b] == j, (w, s, b, W.row, j)
                p = (k + j) & 1
                cols = np.arange(p, wd, 2)
                rb = ((k + 1) >> 1) & 1
                if b == 1:
                    assert W.row[0] == j - 1
                    ut, vt = W.du[0][cols], W.dv[0][cols]
                    nb = pub[(j + 1) % WIN][rb][p]
                    ub, vb = (nb[0], nb[1]) if j + 1 < n else (0 * nb[0], 0 * nb[1])
                else:
                    nb = pub[(j - 1) % WIN][rb][p]
                    ut, vt = (nb[0], nb[1]) if j - 1 >= 0 else (0 * nb[0], 0 * nb[1])
                    if j + 1 < n:
                        assert W.row[1] == j + 1, (w, s, W.row, j)
                        ub, vb = W.du[1][cols], W.dv[1][cols]
                    else:
                        ub, vb = 0 * ut, 0 * vt
                du, dv = W.du[b], W.dv[b]
                def nbr(d, off):
                    o = np.zeros(len(cols))
                    c2 = cols + off
                    ok = (c2 >= 0) & (c2 < wd)
                    o[ok] = d[c2[ok]]
                    return o
                psl = np.where(cols > 0, co['ph'][j][np.maximum(cols - 1, 0)], 0.0)
                pst = co['pv'][j - 1][cols] if j > 0 else 0.0
                B1 = psl * nbr(du, -1) + co['ph'][j][cols] * nbr(du, 1) + pst * ut + co['pv'][j][cols] * ub + co['b1'][j][cols]
                B2 = psl * nbr(dv, -1) + co['ph'][j][cols] * nbr(dv, 1) + pst * vt + co['pv'][j][cols] * vb + co['b2'][j][cols]
                nu = du[cols] + omega * (co['a11'][j][cols] * B1 + co['a12'][j][cols] * B2 - du[cols])
                nv = dv[cols] + omega * (co['a12'][j][cols] * B1 + co['a22'][j][cols] * B2 - dv[cols])
                du[cols], dv[cols] = nu, nv
                wb = ((k >> 1) + 1) & 1
                pub[j % WIN][wb][p] = (nu.copy(), nv.copy())
        else:
            jo, jn = s - 17, s + 1
            if 0 <= jo < n:
                assert W.row[b] == jo
                out_u[jo], out_v[jo] = W.du[b], W.dv[b]
            if 0 <= jn < n:
                W.row[b] = jn
                W.du[b], W.dv[b] = du0[jn].copy(), dv0[jn].copy()
                for p in range(2):
                    pub[jn % WIN][0][p] = (W.du[b][p::2].copy(), W.dv[b][p::2].copy())
        cnt[w] = s
        W.s += 1

    live = list(warps)
    while live:
        cand = [W for W in live if ready(W)]
        assert cand, "deadlock"
        W = cand[rng.randint(len(cand))]
        turn(W)
        if W.s > last:
            live.remove(W)
    hy = nlev
    a, b = hy, n - hy
    eu = np.abs(out_u[a:b] - ref_u[a:b]).max()
    ev = np.abs(out_v[a:b] - ref_v[a:b]).max()
    full = np.abs(out_u - ref_u).max()
    return eu, ev, full


if __name__ == "__main__":
    for (n, wd, nlev) in [(40, 16, 8), (61, 10, 8), (19, 7, 6), (128, 12, 8), (30, 9, 4), (50, 8, 2)]:
        for seed in range(3):
            eu, ev, full = run(n, wd, nlev, seed)
            print(n, wd, nlev, seed, "interior err", eu, ev, "whole-domain err", full)
            assert eu == 0 and ev == 0
    print("ok")
