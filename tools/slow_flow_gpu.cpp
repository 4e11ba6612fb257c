// slow_flow_gpu -- the window loop of slow_flow.cpp:706-1030 (deep_matching = 0 path) on the GPUs of one box, written
// against the reference-shaped C++ shim (include/variational_mt_gpu.hpp) and the C ABI only: no CUDA, no OpenCV.
//
//   slow_flow_gpu --frames 'seq/frame_%d.ppm' --start 10 --jets 4 --out out/ [--S 3] [--skip 1] [--gpus N]
//                 [--threads-per-gpu T] [--no-frame-cache] [--scale 0.5] [--occlusions] [--set key=value ...]
//
// What it reproduces of the reference (and what it does not):
//   * frame indexing: frames = 1 + (Jets + 2)*steps images with index start - ref*skip + k*skip (slow_flow.cpp:411, 446-450);
//     jet j: forward window &seq[j*steps], backward window the reversed frames starting at j*steps + 3*steps (:710-724)
//   * per frame: optional pre-scale (GaussianBlur + resize, :538-542); normalize() over the loaded frames (:673)
//   * per jet: zero initial flow (:865-868), Variational_MT forward (with all-one channel weights, :596-598, 875-888) and
//     backward (:1018-1023), flow x steps (:908-909, :1026-1027)
//   * outputs: <out>/<name>.flo, <out>/<name>_back.flo (:789, :953; io.c:78-96), <out>/occlusion/frame_%i.pbm (:893-905),
//     <out>/config.cfg (:684-688; key <tab> value lines, the keys dense_tracking.cpp reads: slow_flow_S, jet_fps, ...)
//   * parallel axis: the reference's `#pragma omp parallel for` over jets (:706) becomes one host thread per GPU, each
//     with its own context on its own device; jets are split into contiguous ranges; there is no inter-GPU traffic
//   * NOT here: image decoding other than binary PPM (P6), DeepMatching / EPIC initialisation, ground-truth evaluation,
//     flow visualisations.
#include <errno.h>
#include <stdint.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <map>
#include <sstream>
#include <string>
#include <algorithm>
#include <thread>
#include <vector>

#include "variational_mt_gpu.hpp"

// the lookups of utils/parameter_list.h:20-143 that the shim uses, over an ordered map
class ParameterList {
public:
    void insert(std::string k, std::string v, bool overwrite = false) {
        if (!m_.count(k)) order_.push_back(k);
        if (overwrite || !m_.count(k)) m_[k] = v;
    }
    bool exists(std::string k) { return m_.count(k) != 0; }
    std::string parameter(const char *k) { return m_[k]; }
    template <typename T> T parameter(std::string k, std::string def) {
        std::istringstream s(exists(k) ? m_[k] : def);
        T v = T();
        s >> v;
        return v;
    }
    const std::vector<std::string> &keys() const { return order_; }
    const std::string &value(const std::string &k) { return m_[k]; }

private:
    std::map<std::string, std::string> m_;
    std::vector<std::string> order_;
};
template <> inline bool ParameterList::parameter<bool>(std::string k, std::string def) { return (exists(k) ? m_[k] : def) != "0"; }

static void die(const std::string &msg) {
    fprintf(stderr, "slow_flow_gpu: %s\n", msg.c_str());
    exit(1);
}

// image_new / color_image_new (image.c:17-33, 71-89): 16-byte aligned, stride = ceil4(width), planar colour
static image_t *new_image(int w, int h) {
    image_t *im = (image_t *)malloc(sizeof(image_t));
    im->width = w; im->height = h; im->stride = (w + 3) / 4 * 4;
    if (posix_memalign((void **)&im->data, 16, sizeof(float) * im->stride * h)) die("out of memory");
    memset(im->data, 0, sizeof(float) * im->stride * h);
    return im;
}
static color_image_t *new_color_image(int w, int h) {
    color_image_t *im = (color_image_t *)malloc(sizeof(color_image_t));
    im->width = w; im->height = h; im->stride = (w + 3) / 4 * 4;
    if (posix_memalign((void **)&im->c1, 16, sizeof(float) * 3 * im->stride * h)) die("out of memory");
    memset(im->c1, 0, sizeof(float) * 3 * im->stride * h);
    im->c2 = im->c1 + im->stride * h;
    im->c3 = im->c2 + im->stride * h;
    return im;
}
static void free_image(image_t *im) { free(im->data); free(im); }
static void free_color_image(color_image_t *im) { free(im->c1); free(im); }

// binary PPM (P6), 8 or 16 bit -> float RGB planes in file units (the reference converts to CV_32F first, :475)
static color_image_t *read_ppm(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) die("cannot open " + path);
    auto token = [&]() {
        std::string t;
        int c;
        for (;;) {
            c = fgetc(f);
            if (c == '#') { while (c != '\n' && c != EOF) c = fgetc(f); continue; }
            if (c == EOF || !isspace(c)) break;
        }
        while (c != EOF && !isspace(c)) { t.push_back((char)c); c = fgetc(f); }
        return t;
    };
    if (token() != "P6") die(path + " is not a binary PPM");
    const int w = atoi(token().c_str()), h = atoi(token().c_str()), maxv = atoi(token().c_str());
    if (w <= 0 || h <= 0 || maxv <= 0 || maxv > 65535) die("bad PPM header in " + path);
    const int bps = maxv > 255 ? 2 : 1;
    std::vector<unsigned char> raw((size_t)w * h * 3 * bps);
    if (fread(raw.data(), 1, raw.size(), f) != raw.size()) die("truncated PPM " + path);
    fclose(f);
    color_image_t *im = new_color_image(w, h);
    float *pl[3] = {im->c1, im->c2, im->c3};
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int c = 0; c < 3; c++) {
                const size_t k = ((size_t)y * w + x) * 3 + c;
                pl[c][(size_t)y * im->stride + x] = bps == 1 ? (float)raw[k] : (float)((raw[2 * k] << 8) | raw[2 * k + 1]);
            }
    return im;
}

static void make_dir(const std::string &d) {
    if (mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) die("cannot create directory " + d);
}

int main(int argc, char **argv) {
    std::string frames_fmt, out;
    int start = 0, jets = 1, skip = 1, gpus = 0, per_gpu = 1;
    float scale = 1.0f;
    bool frame_cache = true; // --no-frame-cache: upload every window's frames like an unmodified caller (A/B)
    bool resume = false; // -resume: a window whose .flo already exists is skipped (slow_flow.cpp:794, :958)
    ParameterList cfg;
    cfg.insert("slow_flow_S", "3", true);
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) die("missing value after " + a); return argv[++i]; };
        if (a == "--frames") frames_fmt = next();
        else if (a == "--out") out = next();
        else if (a == "--start") start = atoi(next().c_str());
        else if (a == "--jets") jets = atoi(next().c_str());
        else if (a == "--skip") skip = atoi(next().c_str());
        else if (a == "--gpus") gpus = atoi(next().c_str());
        else if (a == "--threads-per-gpu") per_gpu = atoi(next().c_str());
        else if (a == "--no-frame-cache") frame_cache = false;
        else if (a == "--scale") scale = (float)atof(next().c_str());
        else if (a == "--S") cfg.insert("slow_flow_S", next(), true);
        else if (a == "--occlusions") cfg.insert("slow_flow_output_occlusions", "1", true);
        else if (a == "--resume" || a == "-resume") resume = true; // slow_flow.cpp:178-179
        else if (a == "--set") {
            const std::string kv = next();
            const size_t eq = kv.find('=');
            if (eq == std::string::npos) die("--set expects key=value");
            cfg.insert(kv.substr(0, eq), kv.substr(eq + 1), true);
        } else die("unknown argument " + a);
    }
    if (frames_fmt.empty() || out.empty() || jets < 1 || skip < 1) die("usage: --frames FMT --out DIR --start N --jets J [...]");
    if (out[out.size() - 1] != '/') out += "/";
    const int n_dev = sfgpu_device_count();
    if (n_dev <= 0) die("no CUDA device (this path has no CPU fallback)");
    if (gpus <= 0 || gpus > n_dev) gpus = n_dev;
    if (gpus > jets) gpus = jets;
    // more than one host thread per device keeps the GPU busy while another window sits in its host-side steps
    // (occlusion min-cut, early-exit read-backs); every thread has its own context and stream
    if (per_gpu < 1) per_gpu = 1;
    int workers = gpus * per_gpu;
    if (workers > jets) workers = jets;

    const int steps = cfg.parameter<int>("slow_flow_S", "2") - 1, ref = steps; // slow_flow.cpp:208-209
    if (steps < 1) die("slow_flow_S must be >= 2");
    const int frames = 1 + (jets + 2) * steps; // :411

    // ---- load (and pre-scale) the frames, normalise them as one sequence
    std::vector<color_image_t *> seq(frames);
    {
        sfgpu_ctx *ctx = NULL;
        if (scale != 1.0f && sfgpu_create(0, NULL, &ctx) != SFGPU_OK) die(sfgpu_last_error());
        for (int k = 0; k < frames; k++) {
            char path[1024];
            snprintf(path, sizeof(path), frames_fmt.c_str(), start - ref * skip + k * skip);
            color_image_t *im = read_ppm(path);
            if (scale != 1.0f) { // :538-542
                int w = 0, h = 0;
                if (sfgpu_prescale_size(im->width, im->height, scale, &w, &h) != SFGPU_OK) die(sfgpu_last_error());
                color_image_t *small = new_color_image(w, h);
                if (sfgpu_prescale(ctx, small, im, scale) != SFGPU_OK) die(sfgpu_last_error());
                free_color_image(im);
                im = small;
            }
            if (k > 0 && (im->width != seq[0]->width || im->height != seq[0]->height)) die("frames differ in size");
            seq[k] = im;
        }
        if (ctx) sfgpu_destroy(ctx);
    }
    const int W = seq[0]->width, H = seq[0]->height;
    normalize(&seq[0], (unsigned)frames, cfg); // :673 (publishes slow_flow_img_norm_* in cfg)
    std::vector<color_image_t *> seq_back(frames);
    for (int k = 0; k < frames; k++) seq_back[frames - 1 - k] = seq[k]; // :590-591
    color_image_t *channel_weights = new_color_image(W, H);            // :596-598
    for (size_t k = 0; k < (size_t)3 * channel_weights->stride * H; k++) channel_weights->c1[k] = 1.0f;

    // The frames and the channel weights are read by every window: page-lock them once, so that the per-window uploads are
    // plain DMA at PCIe speed instead of chunked staging through host threads (sf_hostcopy.cu)
    if (sfgpu_set_device(0) == SFGPU_OK) {
        for (int k = 0; k < frames; k++) sfgpu_host_register(seq[k]->c1, sizeof(float) * 3 * (unsigned long long)seq[k]->stride * H);
        sfgpu_host_register(channel_weights->c1, sizeof(float) * 3 * (unsigned long long)channel_weights->stride * H);
    }

    make_dir(out);
    const bool write_occ = cfg.parameter<bool>("slow_flow_output_occlusions", "0");
    if (write_occ) make_dir(out + "occlusion");
    // flow file name pattern: the frame pattern's base name without directory and extension (:242)
    std::string name = frames_fmt.substr(frames_fmt.find_last_of('/') == std::string::npos ? 0 : frames_fmt.find_last_of('/') + 1);
    name = name.substr(0, name.find_last_of('.'));

    // ---- one host thread per device, contiguous jet ranges (the reference's omp parallel for over jets, :706)
    const auto t_loop = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    std::vector<int> done(workers, 0), skipped(workers, 0);
    std::vector<double> t_solve(workers, 0.0), t_out(workers, 0.0), t_setup(workers, 0.0), t_total(workers, 0.0), t_end(workers, 0.0); // seconds inside variational() / writing results, per worker
    for (int t = 0; t < workers; t++) {
        pool.emplace_back([&, t]() {
            const auto t_w0 = std::chrono::steady_clock::now();
            if (sfgpu_set_device(t % gpus) != SFGPU_OK) die(sfgpu_last_error());
            // the frames do not change any more (normalize() ran above): keep the ones this worker's windows share on the
            // device -- the backward window of a jet reads the frames of its forward window, the next jet shares all but
            // `steps` of them -- instead of uploading 2 * steps + 1 frames per window
            if (frame_cache && sfgpu_mt_frame_cache(sf_shim_thread_context(), 2 * (2 * steps + 1)) != SFGPU_OK) die(sfgpu_last_error());
            t_setup[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_w0).count();
            const int base = jets / workers, extra = jets % workers;
            const int lo = t * base + (t < extra ? t : extra), hi = lo + base + (t < extra ? 1 : 0);
            Variational_MT minimzer_f, minimzer_b; // contexts are created on this thread's device at first use
            minimzer_f.setChannelWeights(channel_weights);
            if (cfg.exists("method") && cfg.parameter("method") == "forward") minimzer_b.one_direction = true; // :1019-1020
            for (int j = lo; j < hi; j++) {
                ParameterList thread_params(cfg);
                const int f = j * steps;
                color_image_t **im = &seq[f];
                color_image_t **im_back = &seq_back[frames - 1 - f - 3 * steps];
                char path[1200];
                snprintf(path, sizeof(path), (out + name + ".flo").c_str(), start + f * skip);
                if (resume && access(path, F_OK) == 0) skipped[t]++; // skip finished frames (slow_flow.cpp:794)
                else {   // forward
                    image_t *wx = new_image(W, H), *wy = new_image(W, H); // zero initial flow (:865-868)
                    const auto t0 = std::chrono::steady_clock::now();
                    minimzer_f.variational(wx, wy, im, thread_params);
                    const auto t1 = std::chrono::steady_clock::now();
                    t_solve[t] += std::chrono::duration<double>(t1 - t0).count();
                    if (write_occ) {
                        snprintf(path, sizeof(path), "%socclusion/frame_%i.pbm", out.c_str(), start + f * skip);
                        if (sfgpu_write_occlusion_pbm(path, minimzer_f.getOcclusions()) != SFGPU_OK) die(sfgpu_last_error());
                    }
                    for (size_t k = 0; k < (size_t)wx->stride * H; k++) { wx->data[k] *= steps; wy->data[k] *= steps; } // :908-909
                    snprintf(path, sizeof(path), (out + name + ".flo").c_str(), start + f * skip);
                    if (sfgpu_write_flo(path, wx, wy) != SFGPU_OK) die(sfgpu_last_error());
                    free_image(wx); free_image(wy);
                    t_out[t] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
                }
                snprintf(path, sizeof(path), (out + name + "_back.flo").c_str(), start + f * skip + steps * skip);
                if (resume && access(path, F_OK) == 0) skipped[t]++; // (slow_flow.cpp:958)
                else {   // backward
                    image_t *wx = new_image(W, H), *wy = new_image(W, H);
                    const auto t0 = std::chrono::steady_clock::now();
                    minimzer_b.variational(wx, wy, im_back, thread_params);
                    const auto t1 = std::chrono::steady_clock::now();
                    t_solve[t] += std::chrono::duration<double>(t1 - t0).count();
                    for (size_t k = 0; k < (size_t)wx->stride * H; k++) { wx->data[k] *= steps; wy->data[k] *= steps; } // :1026-1027
                    snprintf(path, sizeof(path), (out + name + "_back.flo").c_str(), start + f * skip + steps * skip);
                    if (sfgpu_write_flo(path, wx, wy) != SFGPU_OK) die(sfgpu_last_error());
                    free_image(wx); free_image(wy);
                    t_out[t] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
                }
                done[t]++;
            }
            t_total[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_w0).count();
            t_end[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop).count();
        });
    }
    for (auto &th : pool) th.join();
    // the loop ends when the last worker has written its last result; tearing the contexts down (thread exit: ~1 GB of
    // workspace per context goes back to the driver) is reported separately
    const double all_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop).count();
    double loop_s = 0.0;
    for (int t = 0; t < workers; t++) loop_s = std::max(loop_s, t_end[t]);
    printf("window loop: %.3f s, %.2f jets/s (forward + backward flow each); context tear-down %.3f s\n", loop_s, jets / loop_s, all_s - loop_s);

    // ---- config.cfg for the next stage (:684-688)
    {
        FILE *f = fopen((out + "config.cfg").c_str(), "w");
        if (!f) die("cannot write config.cfg");
        fprintf(f, "# SlowFlow variational estimation\n");
        fprintf(f, "file\t\t%s\noutput\t\t%s\n\nstart\t\t%d\nF\t\t%d\nJets\t\t%d\njet_fps\t\t%d\njet_S\t\t%d\n\n", frames_fmt.c_str(), out.c_str(), start,
                frames, jets, steps * skip, steps + 1);
        for (const std::string &k : cfg.keys()) fprintf(f, "%s\t%s\n", k.c_str(), cfg.value(k).c_str());
        fclose(f);
    }
    int total = 0;
    for (int t = 0; t < workers; t++) {
        printf("worker %d (device %d): %d jets, %d finished flow files skipped; %.3f s in variational(), %.3f s scaling + writing results, %.3f s device / context set-up, %.3f s in the loop body\n", t, t % gpus,
               done[t], skipped[t], t_solve[t], t_out[t], t_setup[t], t_total[t]);
        total += done[t];
    }
    printf("%d jets, %dx%d, S=%d, %d device(s), %d host thread(s) per device\n", total, W, H, steps + 1, gpus, per_gpu);
    for (auto im : seq) { sfgpu_host_unregister(im->c1); free_color_image(im); } // (unregistering a buffer that was not registered only returns an error code)
    sfgpu_host_unregister(channel_weights->c1);
    free_color_image(channel_weights);
    return total == jets ? 0 : 1;
}
