// Stand-alone timing harness for k_sor_tiled (debug aid, not part of the library): includes the kernel source,
// fills a 2560x1440 SOR arena with a stable synthetic system and times launch_sor for several fuse factors.
// Build variants with -D flags (see tools/build_tools.sh); with -DSF_SOR_CLOCKS the kernel also reports the
// clocks warp 0 spends per phase.
#include "../slowflow_b200/csrc/sf_sor.cu"
#include <vector>
#include <stdlib.h>
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
    float *d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    // second arena so that consecutive calls do not find the first one in L2 (2 x 162 MB > 126 MB)
    float *d2; cudaMalloc(&d2, h.size() * 4); cudaMemcpy(d2, d, h.size() * 4, cudaMemcpyDeviceToDevice);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    sf::SorPlan plan[2];
    if (!sf::sor_device_init()) return 1;
    if (!sf::sor_plan_init(plan[0], g, d, sms) || !sf::sor_plan_init(plan[1], g, d2, sms)) return 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int fuse = 2; fuse <= 7; fuse++) {
        if (only && fuse != only) continue;
        float best = 1e9; int launches = 0;
        for (int rep = 0; rep < reps; rep++) {
            int cur = 0;
#ifdef SF_SOR_CLOCKS
            unsigned long long z[8] = {0}; cudaMemcpyToSymbol(sf::g_sor_clk, z, sizeof(z));
#endif
            cudaEventRecord(e0);
            launches = sf::launch_sor(0, plan[rep & 1], 30, 1.9f, 0, fuse, &cur, true);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if ((rep > 0 || reps == 1) && ms < best) best = ms;
        }
        printf("fuse %d: 30 sweeps in %.1f us (%d launches, %.1f us/launch, %.2f us/sweep) err=%s\n", fuse, best * 1e3, launches, best * 1e3 / launches,
               best * 1e3 / 30, cudaGetErrorString(cudaGetLastError()));
#ifdef SF_SOR_CLOCKS
        unsigned long long c[8]; cudaMemcpyFromSymbol(c, sf::g_sor_clk, sizeof(c));
        const double n = (double)c[7];
        if (n > 0) printf("   per tile (warp 0, avg over %.0f tiles): tma-wait %.0f  load %.0f  sweeps %.0f (of which barrier wait %.0f)  store %.0f  total %.0f clk\n", n, c[0] / n,
               c[1] / n, c[2] / n, c[3] / n, c[4] / n, (c[0] + c[1] + c[2] + c[4]) / n);
#endif
    }
    // checksum so that the work cannot be elided and variants can be compared
    std::vector<float> out(P); cudaMemcpy(out.data(), d + sf::SP_DUA * P, P * 4, cudaMemcpyDeviceToHost);
    double cs = 0; for (size_t i = 0; i < P; i++) cs += out[i] * (double)((i % 97) + 1);
    printf("checksum %.6f\n", cs);
    return 0;
}
