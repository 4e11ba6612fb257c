// sf_context.cuh -- the per-thread / per-device context behind sfgpu_ctx.
#pragma once
#include <vector>

#include "sf_internal.cuh"

namespace sf {
struct MtWork; // multi-frame workspace (sf_mt.cu), kept between calls
void mt_work_free(MtWork *w);
struct EpicArena; // workspace of the EPIC interpolation (sf_epic.cu)
void epic_arena_free(EpicArena *a);
struct DeviceCut; // device-side grid min-cut workspace (sf_mincut.cu)
void device_cut_free(DeviceCut *d);
struct HostStager; // pinned bounce buffers for pageable caller memory (sf_hostcopy.cu)
void host_stager_free(HostStager *h);
struct HostCopy {
    void *dev;
    void *host;
    size_t bytes;
};
} // namespace sf

struct sfgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;

    int sor_variant = 0; // 0 tiled, 1 per-half-sweep launches
    int data_variant = 0; // 0 fused marching warp+data-term kernel, 1 separate warp + tile kernel (env SLOWFLOW_GPU_DATA_VARIANT)
    int sor_fuse = 0;    // 0 = auto
    int mt_warp_variant = 0; // 0: fused warp + per-frame derivatives (sf_wderivs.cu), 1: k_warp + k_data_term<DK_DERIVS> (env SLOWFLOW_GPU_MT_WARP_VARIANT)
    bool mt_terms_scalar = false; // env SLOWFLOW_GPU_MT_TERMS_SCALAR: one column per thread in the all-terms pass (A/B switch)
    int mt_data_variant = 0; // 0: per-frame derivative planes + one pointwise all-terms pass, 1: one fused kernel per term (env SLOWFLOW_GPU_MT_DATA_VARIANT)

    // ---- level workspace (one allocation, re-made when the geometry grows)
    sf::Geom g{0, 0, 0};
    float *ws = nullptr;
    size_t ws_floats = 0;
    sf::SorPlan sor;
    float *wim = nullptr;   // 3 planes: warped second image (two-frame)
    float *mask = nullptr;  // 1 plane
    float *uu = nullptr, *vv = nullptr;
    float *dpsis = nullptr;

    // ---- host-API staging (device copies of caller buffers)
    float *io = nullptr; // im1(3) im2(3) wx wy   [+ ring for the sequence API]
    size_t io_floats = 0;

    // ---- sequence pipeline (sf_sequence.cu)
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t seq_ev[3][4] = {}; // rings of "uploaded" / "solved" / "downloaded" events, created at first use
    unsigned char *seq_raw = nullptr; // 3 slots of packed 8-/16-bit frames awaiting their on-device conversion
    size_t seq_raw_bytes = 0;        // bytes per slot
    sf::Geom seq_geom{0, 0, 0};      // geometry of the frame ring left by the last sequence call
    int seq_last_slot = -1;          // ring slot that holds the last frame of that call (-1: none)

    // ---- multi-frame frame cache (sfgpu_mt_frame_cache): device copies of level-0 frames keyed by the caller's host
    // pointer, for callers that solve many windows over one immutable sequence (consecutive windows share frames, the
    // backward window of a jet reads the frames of its forward window)
    struct MtFrameSlot { const void *host = nullptr; size_t floats = 0; float *dev = nullptr; unsigned long long stamp = 0; };
    std::vector<MtFrameSlot> mt_cache;
    float *mt_cache_pool = nullptr; // ONE allocation for all slots (cudaMalloc / cudaFree per frame cost tens of ms each next to registered host memory)
    size_t mt_cache_slot_floats = 0;
    unsigned long long mt_cache_clock = 0;
    unsigned long long mt_cache_hits = 0, mt_cache_misses = 0;

    // ---- profiling
    bool prof = false;
    struct EvPair { cudaEvent_t a, b; int kind; };
    std::vector<EvPair> ev_pending;
    std::vector<cudaEvent_t> ev_free;
    sfgpu_profile_t prof_acc{};

    sfgpu_mt_stats_t mt_stats{};
    sf::MtWork *mtw = nullptr;
    sf::HostStager *stager = nullptr;
    sf::DeviceCut *cut = nullptr;
    sf::EpicArena *epic_arena = nullptr;
    bool host_mincut = false; // env SLOWFLOW_GPU_HOST_MINCUT=1: occlusion labelling on the host (sf_gridcut.hpp), A/B only
    bool staged_host_copies = true; // env SLOWFLOW_GPU_STAGED_COPIES=0 switches the multi-threaded staging off

    // helpers
    int ensure_workspace(sf::Geom geom);
    int ensure_io(size_t floats);
    cudaEvent_t get_event();
    void prof_begin(int kind, cudaEvent_t &a);
    void prof_end(int kind, cudaEvent_t a);
    int prof_collect();
};

namespace sf {
// copies between caller (host) buffers and the device for one call, one direction (sf_hostcopy.cu)
int host_copies(sfgpu_ctx *c, const std::vector<HostCopy> &list, bool h2d);
// two-frame refinement on device planes (variational.c:19-82 + :101-143)
int run_two_frame(sfgpu_ctx *c, Geom g, float *d_wx, float *d_wy, const float *d_im1, const float *d_im2,
                  const variational_params_t *params);
// argument check shared by the host-buffer entries: one geometry, stride = ceil4(width), planar contiguous colour
bool check_pair(const image_t *wx, const image_t *wy, const color_image_t *im1, const color_image_t *im2);
bool is_pageable(const void *p);
// exact binary Potts min-cut on the device (sf_mincut.cu); results in occ / device_cut_labels(), status after a stream sync
int device_grid_mincut(sfgpu_ctx *c, DeviceCut *&cut, int W, int H, long long *tr_dev, long long pair_cap, float *occ, int S);
const int *device_cut_status(const DeviceCut *cut);       // {0 = converged, phases, grid-wide passes}
const unsigned char *device_cut_labels(const DeviceCut *cut); // dense W*H labels (device memory)
// one sor_coupled call on the context's SOR arena (profiled)
int run_sor(sfgpu_ctx *c, int iterations, float omega, int *cur, bool zero_init);
} // namespace sf
