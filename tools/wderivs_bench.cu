// Stand-alone timing harness for k_warp_derivs (debug aid, not part of the library): the fused multi-frame warp +
// per-frame derivative kernel against the two kernels it replaces, for several segment lengths / residencies.
#include "../slowflow_b200/csrc/sf_wderivs.cu"
#include "../slowflow_b200/csrc/sf_data.cu"
#include "../slowflow_b200/csrc/sf_kernels.cu"
#include <math.h>
#include <stdlib.h>
#include <vector>
namespace sf {
void set_error(const std::string &m) { fprintf(stderr, "error: %s\n", m.c_str()); }
bool cuda_ok(cudaError_t e, const char *what) { if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return false; } return true; }
Penalty make_penalty(int type, float eps, float trunc) { Penalty p; p.type = type; p.eps_sq_f = eps * eps; p.eps_sq_d = (double)eps * eps; p.trunc = trunc; return p; }
}
int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 1280, H = argc > 2 ? atoi(argv[2]) : 1024, reps = argc > 3 ? atoi(argv[3]) : 40;
    sf::Geom g = sf::make_geom(W, H);
    const size_t P = g.plane();
    std::vector<float> h(5 * P, 0.f); // frame(3) wx wy
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t o = (size_t)y * g.S + x;
            for (int c = 0; c < 3; c++) h[c * P + o] = 127.f + 60.f * sinf(0.11f * x + 0.07f * y + c) + 30.f * sinf(0.53f * x - 0.31f * y + 2 * c);
            h[3 * P + o] = 1.3f + 0.4f * sinf(0.013f * y) + ((x * 7 + y * 13) % 17) * 0.02f;
            h[4 * P + o] = 0.6f + 0.3f * cosf(0.017f * x) + ((x * 5 + y * 11) % 13) * 0.02f;
        }
    float *d[2];
    for (int k = 0; k < 2; k++) { cudaMalloc(&d[k], (5 + 19) * P * 4); cudaMemcpy(d[k], h.data(), h.size() * 4, cudaMemcpyHostToDevice); }
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    sf::data_term_device_init();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time_it = [&](const char *name, auto run) {
        for (int k = 0; k < 4; k++) run(d[k & 1]);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int k = 0; k < reps; k++) run(d[k & 1]);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<float> out(19 * P); cudaMemcpy(out.data(), d[0] + 5 * P, 19 * P * 4, cudaMemcpyDeviceToHost);
        double cs = 0; for (size_t i = 0; i < 19 * P; i++) cs += out[i] * (double)((i % 97) + 1);
        printf("%-34s %dx%d: %7.2f us per frame  (96 B/px -> %.0f GB/s)  checksum %.6e err=%s\n", name, W, H, ms * 1e3 / reps,
               96.0 * W * H / (ms * 1e-3 / reps) / 1e9, cs, cudaGetErrorString(cudaGetLastError()));
    };
    time_it("k_warp + k_data_term<DERIVS>", [&](float *b) {
        sf::launch_warp(0, g, b, b + 3 * P, b + 4 * P, 2, b + 5 * P, b + 8 * P);
        sf::launch_frame_derivs(0, g, b + 5 * P, b + 9 * P);
    });
    const int rows[] = {4, 6, 8, 12};
    for (int r : rows) {
        sf::g_wd_min_rows = r;
        char name[64]; snprintf(name, sizeof(name), "k_warp_derivs, >= %d rows", r);
        time_it(name, [&](float *b) { sf::launch_warp_derivs(0, g, sms, b, b + 3 * P, b + 4 * P, 2, b + 5 * P, b + 8 * P, b + 9 * P); });
    }
    return 0;
}
