// sf_sequence.cu -- pipelined refinement of consecutive frame pairs through host buffers (BASELINE config 5; the shard
// loop of adaptiveFR.cpp:496-574 / slow_flow.cpp:706): sfgpu_variational_sequence (fp32 planar frames) and
// sfgpu_variational_sequence_u8 / _u16 (the integer images the reference holds before its float conversion,
// adaptiveFR.cpp:450-464 colorMat2colorImg<Vec3b> / mat2colorImg<uchar>, slow_flow.cpp:470-477 convertTo(CV_32F)).
//
// Three streams and events only: uploads of pair j+1 and downloads of pair j-1 overlap the solve of pair j.  A frame
// shared by two pairs is uploaded once (3-slot device ring); with `continue_from_previous` the first frame of a call is
// the last frame of the previous call and is not uploaded at all.  Integer frames cross PCIe as they are (1 or 2 bytes
// per sample, interleaved) and are converted to the planar fp32 layout of image.c:71-89 on the device -- exact, since
// every 8-/16-bit integer is a float -- which takes the upload per 2560x1440 field from 73.7 MB to 40.6 MB.
#include <algorithm>

#include "sf_context.cuh"

namespace sf {

// interleaved (or single-channel) integer frame -> planar fp32; the stride padding is zeroed.  One thread = 4 pixels.
template <typename T>
__global__ void __launch_bounds__(256) k_unpack_frame(Geom g, const unsigned char *__restrict__ raw, size_t step, int channels,
                                                      float *__restrict__ dst) {
    pdl_enter();
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= g.S) return;
    const T *row = reinterpret_cast<const T *>(raw + (size_t)y * step);
    const size_t P = g.plane();
    float4 o[3];
    float *f = reinterpret_cast<float *>(o);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x4 + k;
        if (x < g.W) {
            if (channels == 1) { // mat2colorImg<T>: the grey value goes to all three channels (utils/utils.h:122-130)
                const float v = (float)row[x];
                f[k] = v; f[4 + k] = v; f[8 + k] = v;
            } else { // colorMat2colorImg<Vec3>: channel c of the Mat -> plane c
                f[k] = (float)row[3 * x]; f[4 + k] = (float)row[3 * x + 1]; f[8 + k] = (float)row[3 * x + 2];
            }
        } else {
            f[k] = 0.f; f[4 + k] = 0.f; f[8 + k] = 0.f;
        }
    }
    const size_t off = (size_t)y * g.S + x4;
#pragma unroll
    for (int c = 0; c < 3; c++) *reinterpret_cast<float4 *>(dst + c * P + off) = o[c];
}

struct SeqFrames { // where the frames of a call come from
    const color_image_t *const *f32 = nullptr; // planar fp32 (3*P floats at c1)
    const sf_frame_int_t *ints = nullptr;      // packed integer frames
    int depth = 0;                             // 8 or 16 for ints
    size_t raw_bytes(int f) const { return ints ? (size_t)ints[f].step * ints[f].height : 0; }
    const void *host(int f) const { return ints ? ints[f].data : (const void *)f32[f]->c1; }
};

static int ensure_seq_events(sfgpu_ctx *c) {
    for (auto &ring : c->seq_ev)
        for (auto &e : ring)
            if (!e) SF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (!c->h2d) SF_CUDA(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    if (!c->d2h) SF_CUDA(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    return SFGPU_OK;
}

static void launch_unpack(sfgpu_ctx *c, Geom g, const SeqFrames &src, int f, const unsigned char *raw, float *dst) {
    const dim3 grid((g.S / 4 + 255) / 256, g.H);
    if (src.depth == 8)
        launch_pdl(k_unpack_frame<unsigned char>, grid, dim3(256), 0, c->stream, g, raw, (size_t)src.ints[f].step, src.ints[f].channels, dst);
    else
        launch_pdl(k_unpack_frame<unsigned short>, grid, dim3(256), 0, c->stream, g, raw, (size_t)src.ints[f].step, src.ints[f].channels, dst);
    c->prof_acc.kernel_launches++;
}

static int run_sequence(sfgpu_ctx *c, int n_pairs, const SeqFrames &src, image_t *const *wx, image_t *const *wy,
                        const variational_params_t *params, bool continue_from_previous) {
    SF_CUDA(cudaSetDevice(c->device));
    const Geom g{wx[0]->width, wx[0]->height, wx[0]->stride};
    const size_t P = g.plane();
    const bool keep = continue_from_previous && c->seq_last_slot >= 0 && c->seq_geom.W == g.W && c->seq_geom.H == g.H;
    if (continue_from_previous && !keep) {
        set_error("sequence: continue_from_previous needs a previous sequence call of the same geometry on this context");
        return SFGPU_ERR_ARG;
    }
    const int base = keep ? c->seq_last_slot : 0; // ring slot of frame 0
    c->seq_last_slot = -1;
    // device ring: 3 frame slots, 3 flow slots (pair j -> slot j%3)
    const float *io_before = c->io;
    int rc = c->ensure_io((3 * 3 + 3 * 2) * P);
    if (rc != SFGPU_OK) return rc;
    if (keep && c->io != io_before) {
        set_error("sequence: the frame ring was re-allocated since the previous call");
        return SFGPU_ERR_ARG;
    }
    rc = c->ensure_workspace(g);
    if (rc != SFGPU_OK) return rc;
    rc = ensure_seq_events(c);
    if (rc != SFGPU_OK) return rc;
    size_t raw_slot = 0;
    if (src.ints) {
        for (int f = 0; f <= n_pairs; f++) raw_slot = std::max(raw_slot, src.raw_bytes(f));
        raw_slot = (raw_slot + 255) & ~(size_t)255;
        if (raw_slot > c->seq_raw_bytes) {
            SF_CUDA(cudaStreamSynchronize(c->stream));
            if (c->seq_raw) cudaFree(c->seq_raw);
            c->seq_raw = nullptr;
            c->seq_raw_bytes = 0;
            SF_CUDA(cudaMalloc(&c->seq_raw, 3 * raw_slot));
            c->seq_raw_bytes = raw_slot;
        }
        raw_slot = c->seq_raw_bytes;
    }
    auto frame_slot = [&](int f) { return c->io + (size_t)((f + base) % 3) * 3 * P; };
    auto raw_of = [&](int f) { return c->seq_raw + (size_t)((f + base) % 3) * raw_slot; };
    auto flow_slot = [&](int j) { return c->io + 9 * P + (size_t)(j % 3) * 2 * P; };
    cudaEvent_t *up = c->seq_ev[0], *done = c->seq_ev[1], *down = c->seq_ev[2]; // rings of 4: index j & 3

    // Pageable caller memory (the reference's image_new is malloc): cudaMemcpyAsync would stage it synchronously through
    // the driver and no transfer would overlap anything.  Such callers take the multi-threaded staged copies of
    // sf_hostcopy.cu pair by pair (frames are still uploaded once); page-locked / registered memory is pipelined.
    bool pageable = false;
    for (int f = 0; f <= n_pairs && !pageable; f++) pageable = is_pageable(src.host(f));
    for (int j = 0; j < n_pairs && !pageable; j++) pageable = is_pageable(wx[j]->data) || is_pageable(wy[j]->data);

    auto frame_copy = [&](int f) -> HostCopy {
        if (src.ints) return HostCopy{raw_of(f), const_cast<void *>(src.host(f)), src.raw_bytes(f)};
        return HostCopy{frame_slot(f), const_cast<void *>(src.host(f)), 3 * P * sizeof(float)};
    };
    if (pageable) {
        for (int j = 0; j < n_pairs; j++) {
            std::vector<HostCopy> upl;
            if (j == 0 && !keep) upl.push_back(frame_copy(0));
            upl.push_back(frame_copy(j + 1));
            upl.push_back(HostCopy{flow_slot(j), wx[j]->data, P * sizeof(float)});
            upl.push_back(HostCopy{flow_slot(j) + P, wy[j]->data, P * sizeof(float)});
            rc = host_copies(c, upl, true);
            if (rc != SFGPU_OK) return rc;
            if (src.ints) {
                if (j == 0 && !keep) launch_unpack(c, g, src, 0, raw_of(0), frame_slot(0));
                launch_unpack(c, g, src, j + 1, raw_of(j + 1), frame_slot(j + 1));
            }
            rc = run_two_frame(c, g, flow_slot(j), flow_slot(j) + P, frame_slot(j), frame_slot(j + 1), params);
            if (rc != SFGPU_OK) return rc;
            rc = host_copies(c, {{flow_slot(j), wx[j]->data, P * sizeof(float)}, {flow_slot(j) + P, wy[j]->data, P * sizeof(float)}}, false);
            if (rc != SFGPU_OK) return rc;
            SF_CUDA(cudaStreamSynchronize(c->stream));
        }
        c->seq_geom = g;
        c->seq_last_slot = (n_pairs + base) % 3;
        return SFGPU_OK;
    }

    int status = SFGPU_OK;
    auto h2d_copy = [&](const HostCopy &hc) { return cuda_ok(cudaMemcpyAsync(hc.dev, hc.host, hc.bytes, cudaMemcpyHostToDevice, c->h2d), "h2d"); };
    auto upload = [&](int j) -> bool { // inputs of pair j: frame j+1 (and frame 0 for j == 0) + initial flow j
        if (j >= 2 && !cuda_ok(cudaStreamWaitEvent(c->h2d, done[(j - 2) & 3], 0), "wait done")) return false; // frame slot (j+1)%3 last read by pair j-2
        if (j >= 3 && !cuda_ok(cudaStreamWaitEvent(c->h2d, down[(j - 3) & 3], 0), "wait down")) return false; // flow slot j%3 last drained for pair j-3
        if (j == 0 && !keep && !h2d_copy(frame_copy(0))) return false;
        if (!h2d_copy(frame_copy(j + 1))) return false;
        if (!h2d_copy(HostCopy{flow_slot(j), wx[j]->data, P * sizeof(float)})) return false;
        if (!h2d_copy(HostCopy{flow_slot(j) + P, wy[j]->data, P * sizeof(float)})) return false;
        return cuda_ok(cudaEventRecord(up[j & 3], c->h2d), "record up");
    };
    auto download = [&](int j) -> bool {
        if (!cuda_ok(cudaStreamWaitEvent(c->d2h, done[j & 3], 0), "wait done")) return false;
        if (!cuda_ok(cudaMemcpyAsync(wx[j]->data, flow_slot(j), P * sizeof(float), cudaMemcpyDeviceToHost, c->d2h), "d2h wx")) return false;
        if (!cuda_ok(cudaMemcpyAsync(wy[j]->data, flow_slot(j) + P, P * sizeof(float), cudaMemcpyDeviceToHost, c->d2h), "d2h wy")) return false;
        return cuda_ok(cudaEventRecord(down[j & 3], c->d2h), "record down");
    };
    // the previous work on the compute stream may still use the io ring (done[3] is free until pair 3 records it)
    if (!cuda_ok(cudaEventRecord(done[3], c->stream), "record") || !cuda_ok(cudaStreamWaitEvent(c->h2d, done[3], 0), "wait")) status = SFGPU_ERR_CUDA;
    if (status == SFGPU_OK && !upload(0)) status = SFGPU_ERR_CUDA;
    for (int j = 0; j < n_pairs && status == SFGPU_OK; j++) {
        if (!cuda_ok(cudaStreamWaitEvent(c->stream, up[j & 3], 0), "wait up")) { status = SFGPU_ERR_CUDA; break; }
        if (src.ints) {
            if (j == 0 && !keep) launch_unpack(c, g, src, 0, raw_of(0), frame_slot(0));
            launch_unpack(c, g, src, j + 1, raw_of(j + 1), frame_slot(j + 1));
        }
        status = run_two_frame(c, g, flow_slot(j), flow_slot(j) + P, frame_slot(j), frame_slot(j + 1), params);
        if (status != SFGPU_OK) break;
        if (!cuda_ok(cudaEventRecord(done[j & 3], c->stream), "record done")) { status = SFGPU_ERR_CUDA; break; }
        if (j + 1 < n_pairs && !upload(j + 1)) { status = SFGPU_ERR_CUDA; break; }
        if (!download(j)) { status = SFGPU_ERR_CUDA; break; }
    }
    cudaStreamSynchronize(c->h2d);
    cudaStreamSynchronize(c->stream);
    if (!cuda_ok(cudaStreamSynchronize(c->d2h), "sync d2h") && status == SFGPU_OK) status = SFGPU_ERR_CUDA;
    if (status == SFGPU_OK) {
        c->seq_geom = g;
        c->seq_last_slot = (n_pairs + base) % 3;
    }
    return status;
}

static bool check_flows(int n_pairs, image_t *const *wx, image_t *const *wy) {
    for (int j = 0; j < n_pairs; j++) {
        if (!wx[j] || !wy[j] || !wx[j]->data || !wy[j]->data) { set_error("sequence: null flow plane"); return false; }
        if (wx[j]->stride != ((wx[j]->width + 3) / 4) * 4) { set_error("stride must be ceil4(width) (image.c:25)"); return false; }
        if (wx[j]->width != wx[0]->width || wx[j]->height != wx[0]->height || wy[j]->width != wx[0]->width ||
            wy[j]->height != wx[0]->height || wy[j]->stride != wx[0]->stride || wx[j]->stride != wx[0]->stride) {
            set_error("sequence: all pairs must share one geometry");
            return false;
        }
    }
    return true;
}

static int sequence_int(sfgpu_ctx *c, int n_pairs, const sf_frame_int_t *frames, int depth, image_t *const *wx, image_t *const *wy,
                        const variational_params_t *params, int continue_from_previous) {
    if (!c || n_pairs < 0 || !frames || !wx || !wy) {
        set_error("sfgpu_variational_sequence_u8/u16: bad argument");
        return SFGPU_ERR_ARG;
    }
    if (n_pairs == 0) return SFGPU_OK;
    if (!check_flows(n_pairs, wx, wy)) return SFGPU_ERR_ARG;
    const size_t sample = depth / 8;
    for (int f = 0; f <= n_pairs; f++) {
        const sf_frame_int_t &q = frames[f];
        if (!q.data || q.width != wx[0]->width || q.height != wx[0]->height || (q.channels != 1 && q.channels != 3) ||
            q.step < (size_t)q.width * q.channels * sample || (q.step % sample) != 0) {
            set_error("sequence: integer frames must match the flow geometry, have 1 or 3 channels and step >= width*channels*bytes");
            return SFGPU_ERR_ARG;
        }
    }
    SeqFrames src;
    src.ints = frames;
    src.depth = depth;
    return run_sequence(c, n_pairs, src, wx, wy, params, continue_from_previous != 0);
}

} // namespace sf

using namespace sf;

extern "C" {

int sfgpu_variational_sequence(sfgpu_ctx *c, int n_pairs, const color_image_t *const *frames, image_t *const *wx,
                               image_t *const *wy, const variational_params_t *params) {
    if (!c || n_pairs < 0 || !frames || !wx || !wy) {
        set_error("sfgpu_variational_sequence: bad argument");
        return SFGPU_ERR_ARG;
    }
    if (n_pairs == 0) return SFGPU_OK;
    for (int j = 0; j < n_pairs; j++)
        if (!check_pair(wx[j], wy[j], frames[j], frames[j + 1])) return SFGPU_ERR_ARG;
    if (!check_flows(n_pairs, wx, wy)) return SFGPU_ERR_ARG;
    SeqFrames src;
    src.f32 = frames;
    return run_sequence(c, n_pairs, src, wx, wy, params, false);
}

int sfgpu_variational_sequence_u8(sfgpu_ctx *c, int n_pairs, const sf_frame_int_t *frames, image_t *const *wx, image_t *const *wy,
                                  const variational_params_t *params, int continue_from_previous) {
    return sequence_int(c, n_pairs, frames, 8, wx, wy, params, continue_from_previous);
}

int sfgpu_variational_sequence_u16(sfgpu_ctx *c, int n_pairs, const sf_frame_int_t *frames, image_t *const *wx, image_t *const *wy,
                                   const variational_params_t *params, int continue_from_previous) {
    return sequence_int(c, n_pairs, frames, 16, wx, wy, params, continue_from_previous);
}

} // extern "C"
