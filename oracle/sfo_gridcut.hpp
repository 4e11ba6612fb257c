// sfo_gridcut.hpp -- exact s-t min-cut on a 4-connected W x H grid (Boykov-Kolmogorov search trees,
// integer capacities).  TEST INFRASTRUCTURE ONLY (oracle side); the product has its own copy of the
// algorithm in slowflow_b200/csrc/sf_gridcut.hpp.
//
// Used for the binary occlusion labelling of Variational_AUX_MT::optimizeOcc
// (variational_aux_mt.cpp:758-887): two labels with a Potts pairwise cost are submodular, so the
// alpha-expansion of the un-vendored gco-v3.0 (README.md:33-34) reaches, in its first pass over label 1
// from the all-zero labelling, the global optimum that one min-cut computes directly.
// Labels are canonical: label 1 <=> the pixel can still reach the sink in the residual graph of a
// maximum flow (the unique minimal sink side), so every exact max-flow algorithm gives the same labels.
// Costs are quantised to integers (x 2^24) so that no floating-point residue can blur that set.
#pragma once
#include <stdint.h>
#include <vector>

namespace sfo {

class GridCut {
public:
    typedef int64_t cap_t;
    GridCut(int w, int h) : W(w), H(h), N(w * h), tr(N, 0), rc((size_t)N * 4, 0), parent(N, P_NONE), sink(N, 0),
                            ts(N, 0), dist(N, 0), next(N, -1), qfirst{-1, -1}, qlast{-1, -1}, time_(0), flow_(0) {}

    // terminal capacities: source->p (paid when p ends on the sink side = label 1), p->sink (label 0)
    void set_terminal(int p, cap_t cap_source, cap_t cap_sink) {
        const cap_t m = cap_source < cap_sink ? cap_source : cap_sink;
        flow_ += m;
        tr[p] = cap_source - cap_sink;
    }
    // symmetric neighbour capacity between p and its right (dir 1) / lower (dir 3) neighbour
    void set_edge_right(int p, cap_t c) { rc[(size_t)p * 4 + 1] = c; rc[(size_t)(p + 1) * 4 + 0] = c; }
    void set_edge_down(int p, cap_t c) { rc[(size_t)p * 4 + 3] = c; rc[(size_t)(p + W) * 4 + 2] = c; }

    cap_t maxflow();
    // 1 <=> sink side (can reach the sink in the residual graph)
    int label(int p) const { return (parent[p] != P_NONE && sink[p]) ? 1 : 0; }

private:
    enum { P_TERMINAL = 4, P_ORPHAN = 5, P_NONE = 6 };
    int W, H, N;
    std::vector<cap_t> tr; // > 0: residual source->p ; < 0: residual p->sink
    std::vector<cap_t> rc; // residual capacity of arc p -> neighbour(dir), dir 0 left, 1 right, 2 up, 3 down
    std::vector<uint8_t> parent, sink;
    std::vector<int> ts, dist, next;
    int qfirst[2], qlast[2];
    std::vector<int> orphans;
    size_t orphan_head = 0;
    int time_;
    cap_t flow_;

    inline int nb(int p, int d) const {
        switch (d) {
        case 0: return (p % W > 0) ? p - 1 : -1;
        case 1: return (p % W < W - 1) ? p + 1 : -1;
        case 2: return (p >= W) ? p - W : -1;
        default: return (p < N - W) ? p + W : -1;
        }
    }
    inline cap_t &cap(int p, int d) { return rc[(size_t)p * 4 + d]; }
    void set_active(int i) {
        if (next[i] != -1) return;
        next[i] = i; // end marker
        if (qlast[1] >= 0) next[qlast[1]] = i; else qfirst[1] = i;
        qlast[1] = i;
    }
    int next_active() {
        for (;;) {
            int i = qfirst[0];
            if (i < 0) {
                qfirst[0] = i = qfirst[1]; qlast[0] = qlast[1];
                qfirst[1] = qlast[1] = -1;
                if (i < 0) return -1;
            }
            if (next[i] == i) qfirst[0] = qlast[0] = -1; else qfirst[0] = next[i];
            next[i] = -1;
            if (parent[i] != P_NONE) return i; // only nodes still in a tree are active
        }
    }
    void push_orphan_front(int i) { parent[i] = P_ORPHAN; orphans.insert(orphans.begin() + orphan_head, i); }
    void push_orphan_back(int i) { parent[i] = P_ORPHAN; orphans.push_back(i); }
    void augment(int i, int d);
    void process_orphan(int i, int is_sink);
};

inline void GridCut::augment(int i_src, int d_mid) {
    // middle arc: i_src (source tree) -> j_snk (sink tree) along direction d_mid
    const int j_snk = nb(i_src, d_mid);
    cap_t bottleneck = cap(i_src, d_mid);
    for (int i = i_src;;) { // source tree: flow runs parent -> child, i.e. along sister of parent arc
        const int pd = parent[i];
        if (pd == P_TERMINAL) { if (tr[i] < bottleneck) bottleneck = tr[i]; break; }
        const int p = nb(i, pd);
        const cap_t r = cap(p, pd ^ 1);
        if (r < bottleneck) bottleneck = r;
        i = p;
    }
    for (int i = j_snk;;) { // sink tree: flow runs child -> parent along the parent arc
        const int pd = parent[i];
        if (pd == P_TERMINAL) { if (-tr[i] < bottleneck) bottleneck = -tr[i]; break; }
        const cap_t r = cap(i, pd);
        if (r < bottleneck) bottleneck = r;
        i = nb(i, pd);
    }
    cap(i_src, d_mid) -= bottleneck;
    cap(j_snk, d_mid ^ 1) += bottleneck;
    for (int i = i_src;;) {
        const int pd = parent[i];
        if (pd == P_TERMINAL) { tr[i] -= bottleneck; if (tr[i] == 0) push_orphan_front(i); break; }
        const int p = nb(i, pd);
        cap(i, pd) += bottleneck;
        cap(p, pd ^ 1) -= bottleneck;
        if (cap(p, pd ^ 1) == 0) push_orphan_front(i);
        i = p;
    }
    for (int i = j_snk;;) {
        const int pd = parent[i];
        if (pd == P_TERMINAL) { tr[i] += bottleneck; if (tr[i] == 0) push_orphan_front(i); break; }
        const int p = nb(i, pd);
        cap(p, pd ^ 1) += bottleneck;
        cap(i, pd) -= bottleneck;
        if (cap(i, pd) == 0) push_orphan_front(i);
        i = p;
    }
    flow_ += bottleneck;
}

inline void GridCut::process_orphan(int i, int is_sink) {
    const int INF_D = 1 << 30;
    int best_d = -1, best_dist = INF_D;
    for (int d = 0; d < 4; d++) {
        const int j = nb(i, d);
        if (j < 0) continue;
        // residual in the direction of flow: source tree j -> i, sink tree i -> j
        const cap_t r = is_sink ? cap(i, d) : cap(j, d ^ 1);
        if (r <= 0 || parent[j] == P_NONE || sink[j] != is_sink) continue;
        // does j's path lead to the terminal?
        int dd = 0, k = j;
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
    parent[i] = P_NONE;
    for (int d = 0; d < 4; d++) {
        const int j = nb(i, d);
        if (j < 0 || parent[j] == P_NONE || sink[j] != is_sink) continue;
        const cap_t r = is_sink ? cap(i, d) : cap(j, d ^ 1);
        if (r > 0) set_active(j);
        const int pd = parent[j];
        if (pd < 4 && nb(j, pd) == i) push_orphan_back(j);
    }
}

inline GridCut::cap_t GridCut::maxflow() {
    for (int i = 0; i < N; i++) {
        if (tr[i] > 0) { sink[i] = 0; parent[i] = P_TERMINAL; dist[i] = 1; ts[i] = 0; set_active(i); }
        else if (tr[i] < 0) { sink[i] = 1; parent[i] = P_TERMINAL; dist[i] = 1; ts[i] = 0; set_active(i); }
        else parent[i] = P_NONE;
    }
    int current = -1;
    for (;;) {
        int i = current;
        if (i >= 0) {
            next[i] = -1;
            if (parent[i] == P_NONE) i = -1;
        }
        if (i < 0) {
            i = next_active();
            if (i < 0) break;
        }
        int found_src = -1, found_dir = -1;
        if (!sink[i]) {
            for (int d = 0; d < 4 && found_src < 0; d++) {
                const int j = nb(i, d);
                if (j < 0 || cap(i, d) <= 0) continue;
                if (parent[j] == P_NONE) {
                    sink[j] = 0; parent[j] = (uint8_t)(d ^ 1); ts[j] = ts[i]; dist[j] = dist[i] + 1; set_active(j);
                } else if (sink[j]) {
                    found_src = i; found_dir = d;
                } else if (ts[j] <= ts[i] && dist[j] > dist[i]) {
                    parent[j] = (uint8_t)(d ^ 1); ts[j] = ts[i]; dist[j] = dist[i] + 1;
                }
            }
        } else {
            for (int d = 0; d < 4 && found_src < 0; d++) {
                const int j = nb(i, d);
                if (j < 0 || cap(j, d ^ 1) <= 0) continue;
                if (parent[j] == P_NONE) {
                    sink[j] = 1; parent[j] = (uint8_t)(d ^ 1); ts[j] = ts[i]; dist[j] = dist[i] + 1; set_active(j);
                } else if (!sink[j]) {
                    found_src = j; found_dir = d ^ 1;
                } else if (ts[j] <= ts[i] && dist[j] > dist[i]) {
                    parent[j] = (uint8_t)(d ^ 1); ts[j] = ts[i]; dist[j] = dist[i] + 1;
                }
            }
        }
        time_++;
        if (found_src >= 0) {
            next[i] = i; // keep i active
            current = i;
            augment(found_src, found_dir);
            while (orphan_head < orphans.size()) {
                const int o = orphans[orphan_head++];
                if (orphan_head > 4096 && orphan_head * 2 > orphans.size()) {
                    orphans.erase(orphans.begin(), orphans.begin() + orphan_head);
                    orphan_head = 0;
                }
                process_orphan(o, sink[o]);
            }
            orphans.clear();
            orphan_head = 0;
        } else {
            current = -1;
        }
    }
    return flow_;
}

} // namespace sfo
