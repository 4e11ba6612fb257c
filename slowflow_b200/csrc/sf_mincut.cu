// sf_mincut.cu -- the labelling step of Variational_AUX_MT::optimizeOcc (variational_aux_mt.cpp:851-881) ON THE DEVICE:
// exact s-t min-cut of the binary Potts energy on the 4-connected W x H grid, one cooperative kernel, no host round trip.
//
// The reference calls gco's alpha-expansion (un-vendored).  Two labels + Potts is submodular, so the optimum is
// one min-cut; its canonical labelling (label 1 <=> the pixel can still reach the sink in the residual graph of a maximum
// flow) is unique, whatever maximum flow an algorithm finds -- that is what makes a parallel solver comparable pixel for
// pixel with the host solver (sf_gridcut.hpp) and the oracle's Boykov-Kolmogorov.
//
// Input per pixel: tr[p] = cap(source -> p) - cap(p -> sink), quantised to integers by k_occ_costs; one neighbour capacity.
// Shape of the instances: the occlusion penalty puts almost every pixel on the source (tr > 0), a few percent want the
// sink.  The flow is therefore pushed the cheap way round -- FROM the sink-side pixels (their sink capacity is the
// "excess") TOWARDS the abundant source capacity: a layered (Dinic-style) push scheme on exact distances.
//   phase:  1. BFS: dist[p] = 1 where source capacity is left (tr > 0), else 1 + min over residual arcs; the pixels
//              without source capacity (tr <= 0) are compacted into a work list (a few percent of the image), and the
//              relaxation sweeps run over that list until nothing changes (distances are then exact);
//           2. push rounds on the frozen distances: a pixel with excess sends along arcs to neighbours that are exactly
//              one layer closer, until a round moves nothing.  An arc is only ever written by its farther end in a round,
//              so the net-flow arrays need no atomics; excess counters do (64-bit atomicAdd).
//         A phase whose first round moves nothing has no pixel with excess and a finite distance: the preflow is maximum.
//   labels: pixels with excess left are the roots; label 1 spreads from them over residual arcs (sweeps over the list).
// Every pass is one grid-wide loop followed by grid.sync(); all loops are bounded, a solve that hits a bound reports
// status 1, which the caller reports as an error (SLOWFLOW_GPU_HOST_MINCUT=1 selects the host solver of sf_gridcut.hpp).
#include <cooperative_groups.h>

#include "sf_context.cuh"

namespace cg = cooperative_groups;

namespace sf {

struct CutArgs {
    int W, H, N;
    long long *tr;       // in: terminal capacities (W*H, dense); modified in place (residuals)
    long long *fr, *fd;  // net flow p -> right / lower neighbour, |f| <= cap
    int *dist;
    int *list;           // compacted indices of the pixels with tr <= 0
    unsigned char *lab;  // out: dense labels 0/1
    long long cap;
    int *ctrl;           // [0..3] rotating change flags, [4] list length, [8..] status out: {state, phases, passes}
    int max_phases, max_passes;
    float *occ;          // optional: occlusion plane (+1 label 1, -1 label 0, 0 in the stride padding)
    int S;
};

constexpr int CUT_INF = 0x3fffffff;

__device__ __forceinline__ long long res_left(const CutArgs &a, int p) { return a.cap + a.fr[p - 1]; }
__device__ __forceinline__ long long res_right(const CutArgs &a, int p) { return a.cap - a.fr[p]; }
__device__ __forceinline__ long long res_up(const CutArgs &a, int p) { return a.cap + a.fd[p - a.W]; }
__device__ __forceinline__ long long res_down(const CutArgs &a, int p) { return a.cap - a.fd[p]; }

__global__ void __launch_bounds__(256) k_grid_mincut(CutArgs a) {
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int W = a.W, H = a.H, N = a.N;
    volatile int *ctrl = a.ctrl;
    int pass = 0; // identical in every thread: index of the grid-wide pass, selects the rotating flag
    // flag of pass k is ctrl[k & 3]; thread 0 clears the flag of pass k+1 during pass k (last read two syncs ago)
    auto begin_pass = [&]() { if (tid == 0) ctrl[(pass + 1) & 3] = 0; };
    auto raise = [&]() { ctrl[pass & 3] = 1; };
    auto end_pass = [&]() -> bool {
        grid.sync();
        const bool f = ctrl[pass & 3] != 0;
        pass++;
        return f;
    };

    for (int p = tid; p < N; p += nth) { a.fr[p] = 0; a.fd[p] = 0; }
    if (tid == 0) { ctrl[0] = 0; ctrl[1] = 0; ctrl[2] = 0; ctrl[3] = 0; }
    grid.sync();

    int state = 1, phase = 0; // state 0: converged
    for (phase = 0; phase < a.max_phases && pass < a.max_passes; phase++) {
        // ---- 1a. distances of the trivial pixels + work list
        if (tid == 0) ctrl[4] = 0;
        grid.sync();
        for (int base = tid - lane; base < N; base += nth) { // warp-uniform trip count (warp-aggregated list append)
            const int p = base + lane;
            bool hard = false;
            if (p < N) {
                hard = a.tr[p] <= 0;
                a.dist[p] = hard ? CUT_INF : 1;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hard);
            if (m) {
                int off = 0;
                if (lane == 0) off = atomicAdd(a.ctrl + 4, __popc(m));
                off = __shfl_sync(0xffffffffu, off, 0);
                if (hard) a.list[off + __popc(m & ((1u << lane) - 1u))] = p;
            }
        }
        grid.sync();
        const int L = ctrl[4];
        // ---- 1b. relaxation sweeps over the list (in place: values only decrease and always are lengths of real paths)
        for (;;) {
            begin_pass();
            bool changed = false;
            for (int k = tid; k < L; k += nth) {
                const int p = a.list[k], x = p % W, y = p / W;
                int nd = CUT_INF;
                if (x > 0 && res_left(a, p) > 0) nd = min(nd, a.dist[p - 1] + 1);
                if (x < W - 1 && res_right(a, p) > 0) nd = min(nd, a.dist[p + 1] + 1);
                if (y > 0 && res_up(a, p) > 0) nd = min(nd, a.dist[p - W] + 1);
                if (y < H - 1 && res_down(a, p) > 0) nd = min(nd, a.dist[p + W] + 1);
                if (nd < a.dist[p]) { a.dist[p] = nd; changed = true; }
            }
            if (changed) raise();
            if (!end_pass() || pass >= a.max_passes) break;
        }
        // ---- 2. push rounds on the frozen distances
        int round = 0;
        bool moved_any = false;
        for (;; round++) {
            begin_pass();
            bool moved = false;
            for (int k = tid; k < L; k += nth) {
                const int p = a.list[k];
                long long e = -a.tr[p];
                const int d = a.dist[p];
                if (e <= 0 || d >= CUT_INF) continue;
                const int x = p % W, y = p / W;
                long long sent = 0;
                if (x > 0 && a.dist[p - 1] == d - 1) {
                    const long long r = res_left(a, p);
                    if (r > 0) { const long long q = e < r ? e : r; a.fr[p - 1] -= q; atomicAdd((unsigned long long *)(a.tr + p - 1), (unsigned long long)(-q)); e -= q; sent += q; }
                }
                if (e > 0 && x < W - 1 && a.dist[p + 1] == d - 1) {
                    const long long r = res_right(a, p);
                    if (r > 0) { const long long q = e < r ? e : r; a.fr[p] += q; atomicAdd((unsigned long long *)(a.tr + p + 1), (unsigned long long)(-q)); e -= q; sent += q; }
                }
                if (e > 0 && y > 0 && a.dist[p - W] == d - 1) {
                    const long long r = res_up(a, p);
                    if (r > 0) { const long long q = e < r ? e : r; a.fd[p - W] -= q; atomicAdd((unsigned long long *)(a.tr + p - W), (unsigned long long)(-q)); e -= q; sent += q; }
                }
                if (e > 0 && y < H - 1 && a.dist[p + W] == d - 1) {
                    const long long r = res_down(a, p);
                    if (r > 0) { const long long q = e < r ? e : r; a.fd[p] += q; atomicAdd((unsigned long long *)(a.tr + p + W), (unsigned long long)(-q)); e -= q; sent += q; }
                }
                if (sent > 0) {
                    atomicAdd((unsigned long long *)(a.tr + p), (unsigned long long)sent);
                    moved = true;
                }
            }
            if (moved) raise();
            const bool any = end_pass();
            moved_any |= any;
            if (!any || pass >= a.max_passes) break;
        }
        if (!moved_any && pass < a.max_passes) { // fresh exact distances and nothing could move: the preflow is maximum
            state = 0;
            break;
        }
    }

    // ---- labels: roots = pixels with excess left; spread over residual arcs root -> ... -> p (all candidates are in the list)
    for (int p = tid; p < N; p += nth) a.lab[p] = (state == 0 && a.tr[p] < 0) ? 1 : 0;
    grid.sync();
    if (state == 0) {
        const int L = ctrl[4]; // the list of the last phase holds every pixel with tr <= 0
        for (;;) {
            begin_pass();
            bool changed = false;
            for (int k = tid; k < L; k += nth) {
                const int p = a.list[k];
                if (a.lab[p]) continue;
                const int x = p % W, y = p / W;
                bool in = false;
                if (x > 0 && a.lab[p - 1] && res_right(a, p - 1) > 0) in = true;               // left neighbour -> p
                if (!in && x < W - 1 && a.lab[p + 1] && res_left(a, p + 1) > 0) in = true;      // right neighbour -> p
                if (!in && y > 0 && a.lab[p - W] && res_down(a, p - W) > 0) in = true;          // upper neighbour -> p
                if (!in && y < H - 1 && a.lab[p + W] && res_up(a, p + W) > 0) in = true;        // lower neighbour -> p
                if (in) { a.lab[p] = 1; changed = true; }
            }
            if (changed) raise();
            if (!end_pass()) break;
            if (pass >= a.max_passes) { state = 1; break; }
        }
    }
    if (a.occ && state == 0) {
        const int S = a.S;
        for (int q = tid; q < S * H; q += nth) {
            const int x = q % S, y = q / S;
            a.occ[q] = (x < W) ? (a.lab[y * W + x] ? 1.0f : -1.0f) : 0.0f;
        }
    }
    if (tid == 0) {
        a.ctrl[8] = state; a.ctrl[9] = phase; a.ctrl[10] = pass;
        if (state) a.ctrl[11] = 1; // sticky: any unconverged solve since the context was created
    }
}

// ------------------------------------------------------------------------------------------ host side
struct DeviceCut {
    long long *fr = nullptr, *fd = nullptr;
    int *dist = nullptr, *list = nullptr, *ctrl = nullptr;
    unsigned char *lab = nullptr;
    int *ctrl_host = nullptr; // pinned
    size_t nodes = 0;
    int blocks = 0;
    ~DeviceCut() { release(); }
    void release() {
        if (fr) cudaFree(fr);
        if (fd) cudaFree(fd);
        if (dist) cudaFree(dist);
        if (list) cudaFree(list);
        if (ctrl) cudaFree(ctrl);
        if (lab) cudaFree(lab);
        if (ctrl_host) cudaFreeHost(ctrl_host);
        fr = fd = nullptr; dist = list = ctrl = nullptr; lab = nullptr; ctrl_host = nullptr;
        nodes = 0;
    }
    int reserve(size_t n, int num_sms) {
        if (!ctrl) {
            SF_CUDA(cudaMalloc(&ctrl, 16 * sizeof(int)));
            SF_CUDA(cudaMemset(ctrl, 0, 16 * sizeof(int)));
            SF_CUDA(cudaMallocHost(&ctrl_host, 16 * sizeof(int)));
            int per_sm = 0;
            SF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_grid_mincut, 256, 0));
            if (per_sm < 1) { set_error("k_grid_mincut cannot be made resident"); return SFGPU_ERR_CUDA; }
            blocks = num_sms * (per_sm < 2 ? per_sm : 2); // few, fat blocks: grid.sync() cost grows with the block count
        }
        if (n > nodes) {
            if (fr) cudaFree(fr);
            if (fd) cudaFree(fd);
            if (dist) cudaFree(dist);
            if (list) cudaFree(list);
            if (lab) cudaFree(lab);
            fr = fd = nullptr; dist = list = nullptr; lab = nullptr; nodes = 0;
            SF_CUDA(cudaMalloc(&fr, n * sizeof(long long)));
            SF_CUDA(cudaMalloc(&fd, n * sizeof(long long)));
            SF_CUDA(cudaMalloc(&dist, n * sizeof(int)));
            SF_CUDA(cudaMalloc(&list, n * sizeof(int)));
            SF_CUDA(cudaMalloc(&lab, n));
            nodes = n;
        }
        return SFGPU_OK;
    }
};
void device_cut_free(DeviceCut *d) { delete d; }

// Queues the solve on `st`.  tr_dev: W*H terminal capacities (consumed).  occ (stride S) and / or the dense labels in
// cut->lab receive the result.  The status words land in cut->ctrl_host once the stream has been synchronised:
// [8] 0 = converged, [9] phases, [10] grid-wide passes.
int device_grid_mincut(sfgpu_ctx *c, DeviceCut *&cut, int W, int H, long long *tr_dev, long long pair_cap, float *occ, int S) {
    if (!cut) cut = new DeviceCut();
    const size_t N = (size_t)W * H;
    int rc = cut->reserve(N, c->num_sms);
    if (rc != SFGPU_OK) return rc;
    CutArgs a;
    a.W = W; a.H = H; a.N = (int)N;
    a.tr = tr_dev; a.fr = cut->fr; a.fd = cut->fd; a.dist = cut->dist; a.list = cut->list; a.lab = cut->lab;
    a.cap = pair_cap;
    a.ctrl = cut->ctrl;
    a.max_phases = 1 << 14;
    a.max_passes = 1 << 17;
    a.occ = occ; a.S = S;
    void *params[] = {&a};
    SF_CUDA(cudaLaunchCooperativeKernel((void *)k_grid_mincut, dim3(cut->blocks), dim3(256), params, 0, c->stream));
    SF_CUDA(cudaMemcpyAsync(cut->ctrl_host, cut->ctrl, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return SFGPU_OK;
}
const int *device_cut_status(const DeviceCut *cut) { return cut->ctrl_host + 8; } // {last state, phases, passes, sticky failure}
const unsigned char *device_cut_labels(const DeviceCut *cut) { return cut->lab; }

} // namespace sf
