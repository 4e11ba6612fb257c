// Stand-alone timing harness for k_sor_tiled (debug aid, not part of the library): includes the kernel source,
// fills a 2560x1440 SOR arena with a stable synthetic system and times launch_sor for several fuse factors.
// Build variants with -D flags (see tools/build_tools.sh); with -DSF_SOR_CLOCKS the kernel also reports the
// clocks warp 0 spends per phase.
#include "../slowflow_b200/csrc/sf_sor.cu"
#include "../slowflow_b200/csrc/sf_sor_stream.cu"
#include <vector>
#include <stdlib.h>
#include <math.h>
#include <string.h>
namespace sf {
void set_error(const std::string &m) { fprintf(stderr, "error: %s\n", m.c_str()); }
bool cuda_ok(cudaError_t e, const char *what) { if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return false; } return true; }
}
int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 2560, H = argc > 2 ? atoi(argv[2]) : 1440;
    const int only = argc > 3 ? atoi(argv[3]) : 0, reps = argc > 4 ? atoi(argv[4]) : 6;
    sf::Geom g = sf::make_geom(W, H);
    const size_t P = g.plane();
    std::vector<float> h(P * sf::SP_COUNT);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
    for (size_t i = 0; i < P; i++) {
        h[sf::SP_A11 * P + i] = 0.2f + 0.05f * rnd(); h[sf::SP_A12 * P + i] = 0.01f * rnd(); h[sf::SP_A22 * P + i] = 0.2f + 0.05f * rnd();
        h[sf::SP_B1 * P + i] = rnd() - 0.5f; h[sf::SP_B2 * P + i] = rnd() - 0.5f;
        h[sf::SP_PH * P + i] = 0.4f + 0.2f * rnd(); h[sf::SP_PV * P + i] = 0.4f + 0.2f * rnd();
    }
    for (int y = 0; y < H; y++) h[sf::SP_PH * P + (size_t)y * g.S + W - 1] = 0.0f;
    for (int x = 0; x < g.S; x++) h[sf::SP_PV * P + (size_t)(H - 1) * g.S + x] = 0.0f;
    for (int y = 0; y < H; y++) for (int x = W; x < g.S; x++) for (int pl = 0; pl < 7; pl++) h[pl * P + (size_t)y * g.S + x] = 0.0f;
    float *d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    // second arena so that consecutive calls do not find the first one in L2 (2 x 162 MB > 126 MB)
    float *d2; cudaMalloc(&d2, h.size() * 4); cudaMemcpy(d2, d, h.size() * 4, cudaMemcpyDeviceToDevice);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    sf::SorPlan plan[2];
    if (!sf::sor_device_init()) return 1;
    if (!sf::sor_plan_init(plan[0], g, d, sms) || !sf::sor_plan_init(plan[1], g, d2, sms)) return 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int variant = argc > 5 ? atoi(argv[5]) : 0;
    int last_cur = 0;
    for (int fuse = 2; fuse <= 7; fuse++) {
        if (only && fuse != only) continue;
        if (variant == 2 && fuse > 4) continue;
        float best = 1e9; int launches = 0;
        for (int rep = 0; rep < reps; rep++) {
            int cur = 0;
#ifdef SF_SOR_CLOCKS
            unsigned long long z[8] = {0}; cudaMemcpyToSymbol(sf::g_sor_clk, z, sizeof(z)); cudaMemcpyToSymbol(sf::g_sor_hclk, z, sizeof(z));
#endif
#ifdef SF_SS_CLOCKS
            { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(sf::g_ss_clk, z, sizeof(z)); }
#endif
            cudaEventRecord(e0);
            launches = sf::launch_sor(0, plan[rep & 1], 30, 1.9f, variant, fuse, &cur, true);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if ((rep > 0 || reps == 1) && ms < best) best = ms;
            if ((rep & 1) == 0) last_cur = cur;
        }
        printf("variant %d fuse %d: 30 sweeps in %.1f us (%d launches, %.1f us/launch, %.2f us/sweep) err=%s\n", variant, fuse, best * 1e3, launches, best * 1e3 / launches,
               best * 1e3 / 30, cudaGetErrorString(cudaGetLastError()));
#ifdef SF_SS_CLOCKS
        if (variant == 2) {
            unsigned long long c[8]; cudaMemcpyFromSymbol(c, sf::g_ss_clk, sizeof(c));
            const double n = (double)c[7];
            printf("   per warp and launch (avg over %.0f warps): flag wait %.0f  relax %.0f  publish+flag %.0f  reload %.0f (TMA wait %.0f)  total %.0f clk\n", n,
                   c[0] / n, c[1] / n, c[2] / n, c[3] / n, c[4] / n, c[6] / n);
        }
#endif
#ifdef SF_SOR_CLOCKS
        unsigned long long c[8]; cudaMemcpyFromSymbol(c, sf::g_sor_clk, sizeof(c));
        const double n = (double)c[7];
        if (n > 0) printf("   per tile (warp 0, avg over %.0f tiles): tma-wait %.0f (group B %.0f, %.0f waits > 500 clk)  load %.0f  sweeps %.0f (of which barrier wait %.0f)  store %.0f  total %.0f clk\n", n, c[0] / n, c[5] / n, (double)c[6],
               c[1] / n, c[2] / n, c[3] / n, c[4] / n, (c[0] + c[1] + c[2] + c[4]) / n);
        unsigned long long hc[8]; cudaMemcpyFromSymbol(hc, sf::g_sor_hclk, sizeof(hc));
        const double hn = (double)hc[7];
        if (hn > 0) printf("   helper warp per round (avg over %.0f): wait-free %.0f  issue-B %.0f  fetch+wait-B %.0f  issue-A+deps %.0f clk; warp 0 arrive -> loader sees freeb %.0f clk\n", hn,
               hc[0] / hn, hc[1] / hn, hc[2] / hn, hc[3] / hn, hc[4] / hn);
#endif
    }
    // cross-check against the per-half-sweep kernel on a fresh copy of the system (all variants run the same FMA chain)
    std::vector<float> out(P), ref(P);
    cudaMemcpy(out.data(), d + (last_cur ? sf::SP_DUB : sf::SP_DUA) * P, P * 4, cudaMemcpyDeviceToHost);
    if (argc > 6) {
        int cur = 0;
        sf::launch_sor(0, plan[1], 30, 1.9f, 1, 1, &cur, true);
        cudaDeviceSynchronize();
        cudaMemcpy(ref.data(), d2 + (cur ? sf::SP_DUB : sf::SP_DUA) * P, P * 4, cudaMemcpyDeviceToHost);
        double mx = 0; size_t bad = 0, where = 0;
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) { const size_t i = (size_t)y * g.S + x; const double e = fabs((double)out[i] - ref[i]); if (e > mx) { mx = e; where = i; } if (e > 1e-5) bad++; }
        printf("vs per-half-sweep kernel: max |diff| %.3g at (%zu,%zu), %zu pixels > 1e-5, err=%s\n", mx, where % g.S, where / g.S, bad, cudaGetErrorString(cudaGetLastError()));
    }
    // checksum so that the work cannot be elided and variants can be compared
    double cs = 0; for (size_t i = 0; i < P; i++) cs += out[i] * (double)((i % 97) + 1);
    printf("checksum %.6f\n", cs);
    return 0;
}
