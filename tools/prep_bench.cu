// Stand-alone timing harness for k_prep_two_frame (debug aid, not part of the library): includes the kernel source, fills a
// frame pair / flow / diffusivities with a smooth synthetic pattern and times launch_prep_two_frame.  Build variants with
// -D flags (tools/build_tools.sh).  Prints the average launch time and a checksum of the five output planes so that
// variants can be compared for identical results.
#include "../slowflow_b200/csrc/sf_prep.cu"
#include <math.h>
#include <stdlib.h>
#include <vector>
namespace sf {
void set_error(const std::string &m) { fprintf(stderr, "error: %s\n", m.c_str()); }
bool cuda_ok(cudaError_t e, const char *what) { if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return false; } return true; }
}
int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 2560, H = argc > 2 ? atoi(argv[2]) : 1440, reps = argc > 3 ? atoi(argv[3]) : 20;
    const float hd = argc > 4 ? (float)atof(argv[4]) : 0.0f;
    sf::Geom g = sf::make_geom(W, H);
    const size_t P = g.plane();
    std::vector<float> h(15 * P, 0.f); // im1(3) im2(3) wx wy ph pv | out(5)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t o = (size_t)y * g.S + x;
            for (int c = 0; c < 3; c++) {
                h[c * P + o] = 127.f + 60.f * sinf(0.11f * x + 0.07f * y + c) + 30.f * sinf(0.53f * x - 0.31f * y + 2 * c);
                h[(3 + c) * P + o] = 127.f + 60.f * sinf(0.11f * (x - 1.3f) + 0.07f * (y - 0.6f) + c) + 30.f * sinf(0.53f * (x - 1.3f) - 0.31f * (y - 0.6f) + 2 * c);
            }
            h[6 * P + o] = 1.3f + 0.4f * sinf(0.013f * y) + ((x * 7 + y * 13) % 17) * 0.02f;
            h[7 * P + o] = 0.6f + 0.3f * cosf(0.017f * x) + ((x * 5 + y * 11) % 13) * 0.02f;
            h[8 * P + o] = (x < W - 1) ? 0.3f + 0.1f * sinf(0.2f * x) : 0.f;
            h[9 * P + o] = (y < H - 1) ? 0.3f + 0.1f * cosf(0.2f * y) : 0.f;
        }
    float *d[2];
    for (int k = 0; k < 2; k++) { cudaMalloc(&d[k], h.size() * 4); cudaMemcpy(d[k], h.data(), h.size() * 4, cudaMemcpyHostToDevice); }
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](float *b) {
        sf::launch_prep_two_frame(0, g, sms, b, b + 3 * P, b + 6 * P, b + 7 * P, nullptr, nullptr, b + 8 * P, b + 9 * P, hd, 0.71f * 0.5f / 3.f,
                                  b + 10 * P, b + 11 * P, b + 12 * P, b + 13 * P, b + 14 * P);
    };
    for (int k = 0; k < 4; k++) run(d[k & 1]);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int k = 0; k < reps; k++) run(d[k & 1]);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<float> out(5 * P); cudaMemcpy(out.data(), d[0] + 10 * P, 5 * P * 4, cudaMemcpyDeviceToHost);
    double cs = 0; size_t bad = 0;
    for (size_t i = 0; i < 5 * P; i++) { if (!isfinite(out[i])) bad++; else cs += out[i] * (double)((i % 97) + 1); }
    printf("%dx%d hd=%.3f: %.2f us per launch (%d launches back to back) = %.0f GB/s algorithmic (52 B/px); checksum %.6e nonfinite %zu err=%s\n", W, H, hd,
           ms * 1e3 / reps, reps, 52.0 * W * H / (ms * 1e-3 / reps) / 1e9, cs, bad, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
