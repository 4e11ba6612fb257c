// sf_gridcut.hpp -- exact s-t min-cut on a 4-connected W x H grid with ONE neighbour capacity (binary Potts), host side
// of the occlusion labelling step (Variational_AUX_MT::optimizeOcc, variational_aux_mt.cpp:851-881).
//
// The reference calls the un-vendored gco-v3.0 alpha-expansion here (README.md:33-34).  With two labels and a Potts
// pairwise cost the energy is submodular, so one min-cut gives the optimum that expansion converges to from the
// all-zero labelling.  The per-pixel terminal capacities are evaluated AND quantised on the GPU
// (k_occ_terminals, sf_mt.cu); only this graph search runs on the host.
//
// Shape of the instances (measured on the synthetic windows): the occlusion penalty makes almost every pixel prefer
// label 0 (terminal capacity towards the source), a few percent sit on the sink.  The search is therefore rooted at the
// sink only: a forest T of pixels with a non-saturated path to the sink grows over residual arcs; touching a pixel that
// still has source capacity closes an augmenting path source -> pixel -> ... -> root -> sink.  Orphans are re-attached
// as in Boykov-Kolmogorov (sink tree only, with their time-stamp / distance heuristics).  Work is proportional to the
// explored forest, not to W*H; per call the O(W*H) part is four memsets (32-bit flows, no time-stamp clearing).
//   * arcs are implicit: f_right[p], f_down[p] = net flow p -> right / lower neighbour in [-cap, cap];
//     residual(p -> q) = cap - f(p -> q).  No capacity arrays to initialise.
//   * labels are canonical: label 1 <=> the pixel can still reach the sink in the residual graph of a maximum flow
//     <=> it is in T at termination.  Capacities are integers (costs x 2^24), so that set is exact.
// The object is kept by the context and reused (buffers grow, never shrink).
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

namespace sf {

class SinkForestCut {
public:
    typedef int64_t cap_t;
    enum : uint8_t { P_LEFT = 0, P_RIGHT = 1, P_UP = 2, P_DOWN = 3, P_TERMINAL = 4, P_ORPHAN = 5, P_NONE = 6 };

    // tr[p] = cap(source -> p) - cap(p -> sink) (> 0: source side preferred); pair = the neighbour capacity.
    // On return in_forest()[p] != P_NONE  <=>  label 1.  tr is modified in place (residual terminal capacities).
    cap_t solve(int w, int h, cap_t *tr_io, cap_t pair_cap) {
        W = w; H = h; N = w * h; tr = tr_io; cap = pair_cap;
        // net flows fit 32 bits whenever the neighbour capacity does (|f| <= cap): half the memory to clear
        narrow = pair_cap >= 0 && pair_cap < ((cap_t)1 << 30);
        if (narrow) {
            fr32.assign((size_t)N, 0);
            fd32.assign((size_t)N, 0);
        } else {
            fr64.assign((size_t)N, 0);
            fd64.assign((size_t)N, 0);
        }
        parent.assign((size_t)N, (uint8_t)P_NONE);
        // time stamps keep counting across calls, so stale stamps of an earlier solve can never equal the current time
        // and the stamp array needs no clearing (it is cleared when it grows or before the counter could wrap)
        if (ts.size() != (size_t)N || time_ > (1 << 30)) {
            ts.assign((size_t)N, 0);
            time_ = 0;
        }
        dist.resize((size_t)N);
        queued.assign((size_t)N, 0);
        active.clear();
        head = 0;
        orphans.clear();
        ohead = 0;
        time_++; // roots are stamped with the time of this solve's start
        flow_ = 0;
        const int t0 = time_;
        for (int p = 0; p < N; p++)
            if (tr[p] < 0) {
                parent[p] = P_TERMINAL;
                dist[p] = 1;
                ts[p] = t0;
                push_active(p);
            }
        for (;;) {
            const int i = pop_active();
            if (i < 0) break;
            grow(i);
        }
        return flow_;
    }
    const uint8_t *in_forest() const { return parent.data(); }
    int label(int p) const { return parent[p] != P_NONE ? 1 : 0; }

private:
    int W = 0, H = 0, N = 0;
    cap_t *tr = nullptr;
    cap_t cap = 0, flow_ = 0;
    bool narrow = true;
    std::vector<int32_t> fr32, fd32; // net flow towards the right / lower neighbour (|f| <= cap < 2^30)
    std::vector<cap_t> fr64, fd64;   // the same for larger capacities
    std::vector<uint8_t> parent; // direction of the parent arc (towards the sink), or a P_* marker
    std::vector<int> ts, dist;   // BK distance-to-terminal cache
    std::vector<uint8_t> queued;
    std::vector<int> active;
    size_t head = 0;
    std::vector<int> orphans;
    size_t ohead = 0;
    int time_ = 0;

    inline int nb(int p, int d) const {
        switch (d) {
        case P_LEFT: return (p % W > 0) ? p - 1 : -1;
        case P_RIGHT: return (p % W < W - 1) ? p + 1 : -1;
        case P_UP: return (p >= W) ? p - W : -1;
        default: return (p < N - W) ? p + W : -1;
        }
    }
    // residual capacity of the arc p -> nb(p, d) (the neighbour must exist)
    inline cap_t res(int p, int d) const {
        if (narrow) {
            switch (d) {
            case P_LEFT: return cap + fr32[p - 1];
            case P_RIGHT: return cap - fr32[p];
            case P_UP: return cap + fd32[p - W];
            default: return cap - fd32[p];
            }
        }
        switch (d) {
        case P_LEFT: return cap + fr64[p - 1];
        case P_RIGHT: return cap - fr64[p];
        case P_UP: return cap + fd64[p - W];
        default: return cap - fd64[p];
        }
    }
    inline void send(int p, int d, cap_t x) { // x units along p -> nb(p, d)
        if (narrow) {
            const int32_t y = (int32_t)x;
            switch (d) {
            case P_LEFT: fr32[p - 1] -= y; break;
            case P_RIGHT: fr32[p] += y; break;
            case P_UP: fd32[p - W] -= y; break;
            default: fd32[p] += y; break;
            }
            return;
        }
        switch (d) {
        case P_LEFT: fr64[p - 1] -= x; break;
        case P_RIGHT: fr64[p] += x; break;
        case P_UP: fd64[p - W] -= x; break;
        default: fd64[p] += x; break;
        }
    }
    void push_active(int i) {
        if (queued[i]) return;
        queued[i] = 1;
        active.push_back(i);
    }
    int pop_active() {
        while (head < active.size()) {
            const int i = active[head++];
            queued[i] = 0;
            if (head > (1u << 16) && head * 2 > active.size()) {
                active.erase(active.begin(), active.begin() + head);
                head = 0;
            }
            if (parent[i] != P_NONE) return i; // only forest members are active
        }
        active.clear();
        head = 0;
        return -1;
    }
    void make_orphan_front(int i) {
        parent[i] = P_ORPHAN;
        orphans.insert(orphans.begin() + ohead, i);
    }
    void make_orphan_back(int i) {
        parent[i] = P_ORPHAN;
        orphans.push_back(i);
    }

    // one augmentation: source -> j -> (arc d) -> i -> ... -> root -> sink
    void augment(int j, int d) {
        const int i0 = nb(j, d);
        cap_t b = tr[j];
        const cap_t mid = res(j, d);
        if (mid < b) b = mid;
        for (int i = i0;;) {
            const int pd = parent[i];
            if (pd == P_TERMINAL) {
                if (-tr[i] < b) b = -tr[i];
                break;
            }
            const cap_t r = res(i, pd);
            if (r < b) b = r;
            i = nb(i, pd);
        }
        tr[j] -= b;
        send(j, d, b);
        for (int i = i0;;) {
            const int pd = parent[i];
            if (pd == P_TERMINAL) {
                tr[i] += b;
                if (tr[i] == 0) make_orphan_front(i);
                break;
            }
            const int up = nb(i, pd);
            send(i, pd, b);
            if (res(i, pd) == 0) make_orphan_front(i);
            i = up;
        }
        flow_ += b;
    }

    void adopt(int i) {
        const int INF_D = 1 << 30;
        int best_d = -1, best_dist = INF_D;
        for (int d = 0; d < 4; d++) {
            const int j = nb(i, d);
            if (j < 0 || parent[j] == P_NONE || res(i, d) <= 0) continue;
            int dd = 0, k = j; // does j still hang on the sink?
            for (;;) {
                if (ts[k] == time_) { dd += dist[k]; break; }
                const int pd = parent[k];
                dd++;
                if (pd == P_TERMINAL) { ts[k] = time_; dist[k] = 1; break; }
                if (pd == P_ORPHAN || pd == P_NONE) { dd = INF_D; break; }
                k = nb(k, pd);
            }
            if (dd < INF_D) {
                if (dd < best_dist) { best_dist = dd; best_d = d; }
                for (k = j; ts[k] != time_; k = nb(k, parent[k])) { ts[k] = time_; dist[k] = dd--; }
            }
        }
        if (best_d >= 0) {
            parent[i] = (uint8_t)best_d;
            ts[i] = time_;
            dist[i] = best_dist + 1;
            return;
        }
        parent[i] = P_NONE; // leaves the forest: neighbours may grab it again, its children lose their path
        for (int d = 0; d < 4; d++) {
            const int j = nb(i, d);
            if (j < 0 || parent[j] == P_NONE) continue;
            if (res(i, d) > 0) push_active(j);
            const int pd = parent[j];
            if (pd < 4 && nb(j, pd) == i) make_orphan_back(j);
        }
    }

    void grow(int i) {
        for (;;) { // stays on i while it keeps closing augmenting paths
            int found = -1, found_d = -1;
            for (int d = 0; d < 4 && found < 0; d++) {
                const int j = nb(i, d);
                if (j < 0) continue;
                const int back = d ^ 1; // arc j -> i
                if (res(j, back) <= 0) continue;
                if (parent[j] == P_NONE) {
                    if (tr[j] > 0) { found = j; found_d = back; }
                    else { parent[j] = (uint8_t)back; ts[j] = ts[i]; dist[j] = dist[i] + 1; push_active(j); }
                } else if (parent[j] != P_ORPHAN && ts[j] <= ts[i] && dist[j] > dist[i]) {
                    parent[j] = (uint8_t)back; ts[j] = ts[i]; dist[j] = dist[i] + 1; // shorter route to the sink
                }
            }
            time_++;
            if (found < 0) return;
            augment(found, found_d);
            while (ohead < orphans.size()) {
                const int o = orphans[ohead++];
                if (ohead > 4096 && ohead * 2 > orphans.size()) {
                    orphans.erase(orphans.begin(), orphans.begin() + ohead);
                    ohead = 0;
                }
                adopt(o);
            }
            orphans.clear();
            ohead = 0;
            if (parent[i] == P_NONE) return; // i itself lost its path
        }
    }
};

} // namespace sf
