// Stand-in for gco-v3.0's GCoptimization.h (README.md:33-34; NOT vendored in the reference tree, version
// named but source absent => parity unpinned at this boundary).  TEST INFRASTRUCTURE ONLY.
//
// Only what Variational_AUX_MT::optimizeOcc uses (variational_aux_mt.cpp:774,851-881): a grid graph with
// per-site data costs, a label-pair smoothness table, expansion() and whatLabel().  With 2 labels and a
// Potts table the energy is submodular and alpha-expansion from the all-zero labelling ends in the global
// optimum, computed here by one exact min-cut (oracle/sfo_gridcut.hpp).
//
// gco's EnergyTermType is `int` in the stock distribution and a float type when rebuilt with
// GCO_ENERGYTYPE changed; the reference feeds it costs < 1 (0.01*E/norm + 0.1*l, Potts 0.1), which an int
// build truncates to 0 (labels then stay 0).  Both behaviours are available: sf_gco_int_terms != 0
// truncates every cost to an integer first.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "../../../sfo_gridcut.hpp"

extern "C" int sf_gco_int_terms;   // defined in oracle_api.cpp
extern "C" int sf_gco_calls;       // number of expansion() calls (statistics)

struct GCException {
    const char *message;
    GCException(const char *m = "gco stand-in error") : message(m) {}
    void Report() { fprintf(stderr, "GCException: %s\n", message); }
};

class GCoptimizationGridGraph {
public:
    typedef int SiteID;
    typedef int LabelID;
    GCoptimizationGridGraph(SiteID width, SiteID height, LabelID num_labels)
        : w_(width), h_(height), labels_(num_labels), data_((size_t)width * height * 2, 0.0), label_((size_t)width * height, 0) {
        if (num_labels != 2) throw GCException("stand-in supports exactly 2 labels");
        smooth_[0][0] = smooth_[0][1] = smooth_[1][0] = smooth_[1][1] = 0.0;
    }
    void setDataCost(SiteID s, LabelID l, double e) { data_[(size_t)s * 2 + l] = quant(e); }
    void setSmoothCost(LabelID l1, LabelID l2, double e) { smooth_[l1][l2] = quant(e); }
    double expansion(int /*max_num_iterations*/ = -1) {
        sf_gco_calls++;
        if (smooth_[0][0] != 0.0 || smooth_[1][1] != 0.0 || smooth_[0][1] != smooth_[1][0])
            throw GCException("stand-in supports Potts smoothness only");
        sfo::GridCut g(w_, h_);
        const double S = 16777216.0; // 2^24
        const int64_t pair = (int64_t)llround(smooth_[0][1] * S);
        for (int y = 0; y < h_; y++)
            for (int x = 0; x < w_; x++) {
                const int p = y * w_ + x;
                g.set_terminal(p, (int64_t)llround(data_[(size_t)p * 2 + 1] * S), (int64_t)llround(data_[(size_t)p * 2 + 0] * S));
                if (x + 1 < w_) g.set_edge_right(p, pair);
                if (y + 1 < h_) g.set_edge_down(p, pair);
            }
        const int64_t f = g.maxflow();
        for (int p = 0; p < w_ * h_; p++) label_[p] = g.label(p);
        return (double)f / S;
    }
    LabelID whatLabel(SiteID s) { return label_[s]; }

private:
    static double quant(double e) { return sf_gco_int_terms ? (double)(int)e : e; }
    int w_, h_, labels_;
    std::vector<double> data_;
    double smooth_[2][2];
    std::vector<int> label_;
};
