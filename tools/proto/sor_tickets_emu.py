#!/usr/bin/env python3
"""Protocol emulation of k_sor_tiled's multi-pass launch (slowflow_b200/csrc/sf_sor.cu) under a random scheduler.

Not the arithmetic -- the ORDERING: tickets from one atomic counter (pass-major, tile-minor), the per-tile completion
flags (generation + passes completed), the loader / compute / publisher roles of a CTA with their mbarrier hand-shakes
(fullA[2], fullB, freeb, stored), group A requested two rounds ahead, group B one round ahead.  Checked for every
interleaving the scheduler happens to produce:
  * no deadlock, whatever the number of CTAs that are resident at a time (a CTA waits only for smaller tickets);
  * RAW: when the iterate of (tile, pass p) is read -- at the request of group B and again when it has landed -- every
    tile of the 3x3 neighbourhood holds the data of pass p-1 in the buffer pass p reads;
  * WAR: nothing overwrites that buffer's neighbourhood before the tile has finished reading (same check at landing);
  * every (tile, pass) is computed exactly once and the counters are re-armed for the next launch.
Used by tests/test_host_logic.py; run directly for a longer random search.
"""
import random
import sys


class Deadlock(Exception):
    pass


def emulate(tiles_x, tiles_y, passes, grid, resident, seed, zero_init=True, watch_deps=True):
    rnd = random.Random(seed)
    ntiles = tiles_x * tiles_y
    total = ntiles * passes
    gen = 1000  # generation of this launch
    sync = {"ticket": 0, "exited": 0, "gen": gen}
    done = [gen] * ntiles  # flags of an earlier launch: <= gen
    # ver[buf][tile]: the pass whose result the interior of `tile` holds in iterate buffer `buf` (pass p reads p & 1)
    ver = [[-1] * ntiles, [None] * ntiles]
    computed = set()

    def nbhd(tile):
        tx, ty = tile % tiles_x, tile // tiles_x
        return [ny * tiles_x + nx for ny in (ty - 1, ty, ty + 1) for nx in (tx - 1, tx, tx + 1)
                if 0 <= nx < tiles_x and 0 <= ny < tiles_y]

    def check_read(t):
        p, tile = divmod(t, ntiles)
        if p == 0 and zero_init:
            return
        for n in nbhd(tile):
            assert ver[p & 1][n] == p - 1, ("RAW/WAR hazard", t, n, ver[p & 1][n])

    class Cta:
        def __init__(self):
            self.fullA = [0, 0]   # completed phases
            self.fullB = 0
            self.freeb = 0        # rounds that have left the staging area
            self.stored = 0       # rounds whose stores are out
            self.tickq = {}
            self.published = 0
            self.rounds = -1      # rounds in all, once known
            self.inflight_b = []  # tickets requested, not yet landed
            self.inflight_a = []
            self.exited = 0       # roles that have returned

    def fetch():
        t = sync["ticket"]
        sync["ticket"] += 1
        return t

    def deps_ready(t):
        p, tile = divmod(t, ntiles)
        return p == 0 or not watch_deps or all(done[n] - (gen + p) >= 0 for n in nbhd(tile))

    # ---- the three roles as generators: `yield cond` blocks until cond() is true
    def loader(c):
        tk = [fetch(), None]
        tk[1] = fetch() if tk[0] < total else total
        k = 0
        if tk[0] < total:
            c.inflight_a.append((0, 0))
            yield lambda: deps_ready(tk[0])
        while True:
            t = tk[k & 1]
            if k > 0:
                yield lambda: c.freeb >= k
            yield lambda: c.published >= k - 3
            if t < total:
                c.tickq[k & 3] = t
                check_read(t)
                c.inflight_b.append((k, t))
            else:
                c.tickq[k & 3] = total
                c.rounds = k
                c.fullB += 1
                break
            tn = tk[(k + 1) & 1]
            tk[k & 1] = fetch() if tn < total else total
            if tn < total:
                yield lambda: c.fullB >= k + 1          # group B of round k has landed
                # buffer (k+1)&1 was used by round k-1: it must have been left
                assert k == 0 or c.freeb >= k, "group A overwrites a buffer in use"
                c.inflight_a.append((k + 1, (k + 1) & 1))
                yield lambda: deps_ready(tn)
            k += 1
        c.exited += 1

    def tma(c):
        # the copy engine: lands requested groups at random times, in any order
        while True:
            yield lambda: c.inflight_a or c.inflight_b or c.exited >= 3
            if c.exited >= 3 and not (c.inflight_a or c.inflight_b):
                return
            if c.inflight_b and (not c.inflight_a or rnd.random() < 0.5):
                k, t = c.inflight_b.pop(0)
                check_read(t)
                c.fullB += 1
            else:
                k, buf = c.inflight_a.pop(0)
                c.fullA[buf] += 1

    def compute(c):
        r = 0
        while True:
            yield lambda: c.fullB >= r + 1
            t = c.tickq[r & 3]
            if t >= total:
                break
            yield lambda: c.fullA[r & 1] >= (r >> 1) + 1
            c.freeb += 1
            yield lambda: True  # the half sweeps
            p, tile = divmod(t, ntiles)
            assert (p, tile) not in computed
            computed.add((p, tile))
            ver[(p + 1) & 1][tile] = p  # (a reader of pass p-1 that is still in flight fails its landing check)
            c.stored += 1
            r += 1
        c.exited += 1

    def publisher(c):
        k = 0
        while True:
            yield lambda: c.stored >= k + 1 or (c.rounds >= 0 and k >= c.rounds)
            if not c.stored >= k + 1:
                break
            t = c.tickq[k & 3]
            p, tile = divmod(t, ntiles)
            done[tile] = gen + p + 1
            c.published = k + 1
            k += 1
        sync["exited"] += 1
        if sync["exited"] == grid:
            sync["ticket"] = 0
            sync["exited"] = 0
            sync["gen"] = gen + passes
        c.exited += 1

    # ---- scheduler: `resident` CTAs at a time, every role a coroutine
    waiting_ctas = grid
    actors = []  # [generator, condition, cta]

    def start_cta():
        c = Cta()
        for role in (loader, compute, publisher, tma):
            g = role(c)
            actors.append([g, next(g), c])

    nres = 0
    while waiting_ctas and nres < resident:
        start_cta()
        waiting_ctas -= 1
        nres += 1
    steps = 0
    while actors:
        runnable = [a for a in actors if a[1]()]
        if not runnable:
            raise Deadlock("no runnable actor: %d actors blocked, ticket %d of %d" % (len(actors), sync["ticket"], total))
        a = rnd.choice(runnable)
        try:
            a[1] = next(a[0])
        except StopIteration:
            actors.remove(a)
            if not any(b[2] is a[2] for b in actors):  # the CTA has left the SM
                if waiting_ctas:
                    start_cta()
                    waiting_ctas -= 1
        steps += 1
        if steps > 5_000_000:
            raise Deadlock("livelock")
    assert len(computed) == total, (len(computed), total)
    assert sync == {"ticket": 0, "exited": 0, "gen": gen + passes}, sync
    assert all(d == gen + passes for d in done), done
    return steps


def search(n, seed0=0):
    rnd = random.Random(seed0)
    for i in range(n):
        tx, ty = rnd.randint(1, 6), rnd.randint(1, 5)
        passes = rnd.randint(1, 7)
        grid = rnd.randint(1, 12)
        resident = rnd.randint(1, grid)
        emulate(tx, ty, passes, grid, resident, seed=rnd.randint(0, 1 << 30), zero_init=bool(i & 1))
    return n


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    print("ok: %d random configurations" % search(n))
