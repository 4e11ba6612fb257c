#!/usr/bin/env python
"""Critical-path model of the streaming SOR kernel's dataflow: turn (w, s) starts when the warp's previous turn has ended
and the flagged dependencies have published; a turn publishes `early` clocks after it starts and ends after `dur`."""
import sys
NW = 9

def run(nsteps, relax, rel_early, rel_total, flaglat, pub_relax=None):
    pub_relax = relax if pub_relax is None else pub_relax
    end = {}   # (w, s) -> end time of the turn (warp free again)
    pub = {}   # (w, s) -> publication time
    for w in range(NW):
        for s in range(-8, 2 * w - 1):
            end[(w, s)] = pub[(w, s)] = 0.0
    for s in range(-1, nsteps):
        # within a step the order of warps does not matter (deps point to earlier steps)
        for w in range(NW):
            if (w, s) in end:
                continue
            b = (s + 1) & 1
            phi = ((s - 1 - 2 * w - b) // 2) % 9
            t0 = end[(w, s - 1)]
            if b == 1:
                t0 = max(t0, pub[((w + 1) % NW, s - 1)] + flaglat)
            else:
                t0 = max(t0, pub[((w - 1) % NW, s - 3)] + flaglat)
            if phi == 8:
                pub[(w, s)] = t0 + rel_early
                end[(w, s)] = t0 + rel_total
            else:
                pub[(w, s)] = t0 + pub_relax
                end[(w, s)] = t0 + relax
    return max(end[(w, nsteps - 1)] for w in range(NW))

if __name__ == "__main__":
    n = 146
    for (relax, e, t, fl) in [(260, 2050, 2050, 60), (260, 300, 2050, 60), (260, 300, 1000, 60), (200, 200, 800, 60), (150, 200, 600, 60), (150, 150, 400, 40), (260, 260, 260, 60)]:
        T = run(n, relax, e, t, fl)
        print("relax %4d reload early %4d total %4d flag %3d -> %7.0f clk = %5.1f us (%.0f clk/step)" % (relax, e, t, fl, T, T / 1920.0, T / n))
