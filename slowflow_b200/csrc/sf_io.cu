// sf_io.cu -- the on-disk contract between the flow stage and dense_tracking, host only:
//   <fmt>.flo / <fmt>_back.flo   Middlebury flow files     writeFlowFile / readFlowFile, epic_flow_extended/io.c:50-96
//   occlusion/frame_%i.pbm       occlusion labels          slow_flow.cpp:893-905 (cv::imwrite, PXM binary)
// plus the two device-selection helpers a "one host thread per device" driver needs without linking the CUDA runtime.
#include <math.h>

#include <algorithm>
#include <vector>

#include "sf_internal.cuh"

using namespace sf;

extern "C" {

int sfgpu_set_device(int device) {
    SF_CUDA(cudaSetDevice(device));
    return SFGPU_OK;
}

int sfgpu_get_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    return dev;
}

static bool plane_ok(const image_t *a) { return a && a->data && a->width > 0 && a->height > 0 && a->stride >= a->width; }

// io.c:78-96: float tag 202021.25, int width, int height, then (u, v) interleaved row by row, valid columns only
int sfgpu_write_flo(const char *filename, const image_t *flowx, const image_t *flowy) {
    if (!filename || !plane_ok(flowx) || !plane_ok(flowy) || flowx->width != flowy->width || flowx->height != flowy->height) {
        set_error("sfgpu_write_flo: bad argument");
        return SFGPU_ERR_ARG;
    }
    FILE *f = fopen(filename, "wb");
    if (!f) {
        set_error(std::string("sfgpu_write_flo: cannot open ") + filename);
        return SFGPU_ERR_ARG;
    }
    const float tag = 202021.25f;
    const int w = flowx->width, h = flowx->height;
    bool ok = fwrite(&tag, sizeof(float), 1, f) == 1 && fwrite(&w, sizeof(int), 1, f) == 1 && fwrite(&h, sizeof(int), 1, f) == 1;
    // interleave bands of rows and hand each band to the C library in one piece (one write per band instead of one per
    // row: the writer is part of every window of the sharded driver)
    const int band = std::max(1, (int)(((size_t)4 << 20) / ((size_t)2 * w * sizeof(float))));
    std::vector<float> buf((size_t)2 * w * band);
    for (int y0 = 0; y0 < h && ok; y0 += band) {
        const int rows = std::min(band, h - y0);
        for (int y = 0; y < rows; y++) {
            const float *u = flowx->data + (size_t)(y0 + y) * flowx->stride, *v = flowy->data + (size_t)(y0 + y) * flowy->stride;
            float *dst = buf.data() + (size_t)2 * w * y;
            for (int x = 0; x < w; x++) {
                dst[2 * x] = u[x];
                dst[2 * x + 1] = v[x];
            }
        }
        const size_t n = (size_t)2 * w * rows;
        ok = fwrite(buf.data(), sizeof(float), n, f) == n;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        set_error(std::string("sfgpu_write_flo: write failed for ") + filename);
        return SFGPU_ERR_ARG;
    }
    return SFGPU_OK;
}

// io.c:50-75: the caller asks for the size first, then passes planes of that geometry
int sfgpu_read_flo_size(const char *filename, int *width, int *height) {
    if (!filename || !width || !height) return SFGPU_ERR_ARG;
    FILE *f = fopen(filename, "rb");
    if (!f) {
        set_error(std::string("sfgpu_read_flo_size: cannot open ") + filename);
        return SFGPU_ERR_ARG;
    }
    float tag = 0.f;
    int w = 0, h = 0;
    const bool ok = fread(&tag, sizeof(float), 1, f) == 1 && fread(&w, sizeof(int), 1, f) == 1 && fread(&h, sizeof(int), 1, f) == 1;
    fclose(f);
    if (!ok || tag != 202021.25f || w <= 0 || h <= 0) {
        set_error(std::string("sfgpu_read_flo_size: not a .flo file: ") + filename);
        return SFGPU_ERR_ARG;
    }
    *width = w;
    *height = h;
    return SFGPU_OK;
}

int sfgpu_read_flo(const char *filename, image_t *flowx, image_t *flowy) {
    int w = 0, h = 0;
    const int rc = sfgpu_read_flo_size(filename, &w, &h);
    if (rc != SFGPU_OK) return rc;
    if (!plane_ok(flowx) || !plane_ok(flowy) || flowx->width != w || flowx->height != h || flowy->width != w || flowy->height != h) {
        set_error("sfgpu_read_flo: planes do not have the file's geometry");
        return SFGPU_ERR_ARG;
    }
    FILE *f = fopen(filename, "rb");
    if (!f) return SFGPU_ERR_ARG;
    fseek(f, 12, SEEK_SET);
    std::vector<float> row((size_t)2 * w);
    bool ok = true;
    for (int y = 0; y < h && ok; y++) {
        ok = fread(row.data(), sizeof(float), row.size(), f) == row.size();
        float *u = flowx->data + (size_t)y * flowx->stride, *v = flowy->data + (size_t)y * flowy->stride;
        for (int x = 0; x < w && ok; x++) {
            u[x] = row[2 * x];
            v[x] = row[2 * x + 1];
        }
    }
    fclose(f);
    if (!ok) {
        set_error(std::string("sfgpu_read_flo: truncated file ") + filename);
        return SFGPU_ERR_ARG;
    }
    return SFGPU_OK;
}

// slow_flow.cpp:893-905: occ in {-1, 0, +1} -> 0.5*(occ + 1) -> convertTo(CV_8UC1, 255) -> imwrite(".pbm", PXM_BINARY = 1).
// cv::imwrite stores a .pbm as a P4 bitmap: one bit per pixel, MSB first, rows padded to bytes, bit = 1 where the
// 8-bit value is 0 (PBM's "black") -- so a set bit marks label -1.  (Pinned to python cv2 4.13, tests/test_io.py;
// OpenCV 2.4 named in README.md:17 is not available offline.)
int sfgpu_write_occlusion_pbm(const char *filename, const image_t *occlusions) {
    if (!filename || !plane_ok(occlusions)) {
        set_error("sfgpu_write_occlusion_pbm: bad argument");
        return SFGPU_ERR_ARG;
    }
    FILE *f = fopen(filename, "wb");
    if (!f) {
        set_error(std::string("sfgpu_write_occlusion_pbm: cannot open ") + filename);
        return SFGPU_ERR_ARG;
    }
    const int w = occlusions->width, h = occlusions->height;
    fprintf(f, "P4\n%d %d\n", w, h);
    const size_t rb = (size_t)(w + 7) / 8;
    std::vector<unsigned char> bits(rb * (size_t)h, (unsigned char)0);
    for (int y = 0; y < h; y++) {
        unsigned char *row = bits.data() + rb * (size_t)y;
        const float *o = occlusions->data + (size_t)y * occlusions->stride;
        // a pixel is black (bit 1) when saturate_cast<uchar>(0.5*(occ+1)*255) is 0: the value rounds half to even and
        // clamps at 0, i.e. it is 0 exactly when the double product is <= 0.5 (branch-free: the labels are data)
        auto black = [](float occ) -> unsigned { return (0.5 * ((double)occ + 1.0) * 255.0 <= 0.5) ? 1u : 0u; };
        int x = 0;
        for (; x + 8 <= w; x += 8) { // eight pixels per byte, most significant bit first
            const float *q = o + x;
            row[x >> 3] = (unsigned char)((black(q[0]) << 7) | (black(q[1]) << 6) | (black(q[2]) << 5) | (black(q[3]) << 4) |
                                          (black(q[4]) << 3) | (black(q[5]) << 2) | (black(q[6]) << 1) | black(q[7]));
        }
        for (; x < w; x++) row[x >> 3] |= (unsigned char)((black(o[x]) << 7) >> (x & 7));
    }
    bool ok = fwrite(bits.data(), 1, bits.size(), f) == bits.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        set_error(std::string("sfgpu_write_occlusion_pbm: write failed for ") + filename);
        return SFGPU_ERR_ARG;
    }
    return SFGPU_OK;
}

} // extern "C"
