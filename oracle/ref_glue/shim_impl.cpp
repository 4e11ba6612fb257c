// shim_impl.cpp -- the few cv:: functions the reference's multi-frame driver links against.
// TEST INFRASTRUCTURE ONLY (compiled into oracle/_ref/libsf_ref.so).
//
// Real arithmetic (restated from OpenCV's documented behaviour, pinned to python cv2 4.13 by
// tests/test_oracle_pin.py::test_shim_blur_resize_match_cv2 in the build container):
//   GaussianBlur(CV_32F, ksize = 0): n = cvRound(8 sigma + 1) | 1 taps, kernel exp(-x^2 / 2 sigma^2) normalised,
//       separable (rows then columns), symmetric-sum evaluation, replicate border.
//   resize(INTER_LINEAR): pixel-centre mapping fx = (dx + 0.5) * (sw / dw) - 0.5, clamp, fp32 lerp.
// Everything else (windows, imwrite, split/merge ...) is only reached under verbose bits the oracle never sets.
#include <opencv2/core.hpp>
#include <opencv2/highgui.hpp>
#include <opencv2/imgproc.hpp>

#include <algorithm>
#include <cmath>

namespace cv {

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY, int /*borderType*/) {
    if (sigmaY <= 0) sigmaY = sigmaX;
    const int cn = src.channels(), R = src.rows, C = src.cols;
    auto kernel = [](int n, double sigma) {
        std::vector<float> k(n);
        const double s2 = -0.5 / (sigma * sigma);
        double sum = 0;
        std::vector<double> t(n);
        for (int i = 0; i < n; i++) { const double x = i - (n - 1) * 0.5; t[i] = std::exp(s2 * x * x); sum += t[i]; }
        for (int i = 0; i < n; i++) k[i] = (float)(t[i] / sum);
        return k;
    };
    const int nx = ksize.width > 0 ? ksize.width : ((int)lrint(sigmaX * 8 + 1) | 1);
    const int ny = ksize.height > 0 ? ksize.height : ((int)lrint(sigmaY * 8 + 1) | 1);
    const std::vector<float> kx = kernel(nx, sigmaX), ky = kernel(ny, sigmaY);
    const int rx = nx / 2, ry = ny / 2;
    Mat in = src.clone();
    std::vector<float> tmp((size_t)R * C * cn);
    const float *s = reinterpret_cast<const float *>(in.data);
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++)
            for (int c = 0; c < cn; c++) {
                float acc = kx[rx] * s[((size_t)i * C + j) * cn + c];
                for (int k = 1; k <= rx; k++)
                    acc += kx[rx + k] * (s[((size_t)i * C + clampi(j - k, 0, C - 1)) * cn + c] + s[((size_t)i * C + clampi(j + k, 0, C - 1)) * cn + c]);
                tmp[((size_t)i * C + j) * cn + c] = acc;
            }
    if (dst.rows != R || dst.cols != C || dst.type() != src.type() || dst.data == src.data) dst.create(R, C, src.type());
    float *d = reinterpret_cast<float *>(dst.data);
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++)
            for (int c = 0; c < cn; c++) {
                float acc = ky[ry] * tmp[((size_t)i * C + j) * cn + c];
                for (int k = 1; k <= ry; k++)
                    acc += ky[ry + k] * (tmp[((size_t)clampi(i - k, 0, R - 1) * C + j) * cn + c] + tmp[((size_t)clampi(i + k, 0, R - 1) * C + j) * cn + c]);
                d[((size_t)i * C + j) * cn + c] = acc;
            }
}

void resize(const Mat &src, Mat &dst, Size dsize, double /*fx*/, double /*fy*/, int /*interpolation*/) {
    const int cn = src.channels(), sr = src.rows, sc = src.cols, dr = dsize.height, dc = dsize.width;
    Mat in = src.clone();
    Mat out(dr, dc, src.type());
    const float *s = reinterpret_cast<const float *>(in.data);
    float *d = reinterpret_cast<float *>(out.data);
    const double scale_x = (double)sc / dc, scale_y = (double)sr / dr;
    std::vector<int> xo(dc), yo(dr);
    std::vector<float> xa(dc), ya(dr);
    for (int x = 0; x < dc; x++) {
        float f = (float)((x + 0.5) * scale_x - 0.5);
        int sx = (int)std::floor(f);
        f -= sx;
        if (sx < 0) { sx = 0; f = 0; }
        if (sx >= sc - 1) { sx = sc - 1; f = 0; }
        xo[x] = sx; xa[x] = f;
    }
    for (int y = 0; y < dr; y++) {
        float f = (float)((y + 0.5) * scale_y - 0.5);
        int sy = (int)std::floor(f);
        f -= sy;
        if (sy < 0) { sy = 0; f = 0; }
        if (sy >= sr - 1) { sy = sr - 1; f = 0; }
        yo[y] = sy; ya[y] = f;
    }
    for (int y = 0; y < dr; y++) {
        const int y0 = yo[y], y1 = std::min(y0 + 1, sr - 1);
        const float b1 = ya[y], b0 = 1.f - b1;
        for (int x = 0; x < dc; x++) {
            const int x0 = xo[x], x1 = std::min(x0 + 1, sc - 1);
            const float a1 = xa[x], a0 = 1.f - a1;
            for (int c = 0; c < cn; c++) {
                const float r0 = s[((size_t)y0 * sc + x0) * cn + c] * a0 + s[((size_t)y0 * sc + x1) * cn + c] * a1;
                const float r1 = s[((size_t)y1 * sc + x0) * cn + c] * a0 + s[((size_t)y1 * sc + x1) * cn + c] * a1;
                d[((size_t)y * dc + x) * cn + c] = r0 * b0 + r1 * b1;
            }
        }
    }
    dst = out;
}

void Mat::convertTo(Mat &dst, int rtype, double alpha, double beta) const {
    Mat out(rows, cols, CV_MAKETYPE(rtype & 7, channels()));
    const size_t n = (size_t)rows * cols * channels();
    for (size_t i = 0; i < n; i++) {
        const double v = (depth() == CV_32F ? reinterpret_cast<const float *>(data)[i] : data[i]) * alpha + beta;
        if (out.depth() == CV_32F) reinterpret_cast<float *>(out.data)[i] = (float)v;
        else out.data[i] = (uchar)std::min(255.0, std::max(0.0, std::nearbyint(v)));
    }
    dst = out;
}

static Mat map_scalar(const Mat &a, double s, int op) {
    Mat out = a.clone();
    const size_t n = (size_t)a.rows * a.cols * a.channels();
    float *p = reinterpret_cast<float *>(out.data);
    for (size_t i = 0; i < n; i++) p[i] = (float)(op == 0 ? p[i] + s : op == 1 ? p[i] - s : op == 2 ? p[i] * s : p[i] / s);
    return out;
}
Mat operator+(const Mat &a, double s) { return map_scalar(a, s, 0); }
Mat operator-(const Mat &a, double s) { return map_scalar(a, s, 1); }
Mat operator*(const Mat &a, double s) { return map_scalar(a, s, 2); }
Mat operator*(double s, const Mat &a) { return map_scalar(a, s, 2); }
Mat operator/(const Mat &a, double s) { return map_scalar(a, s, 3); }

void minMaxLoc(const Mat &src, double *mn, double *mx, Point *, Point *) {
    const size_t n = (size_t)src.rows * src.cols * src.channels();
    double lo = 1e300, hi = -1e300;
    for (size_t i = 0; i < n; i++) {
        const double v = src.depth() == CV_32F ? reinterpret_cast<const float *>(src.data)[i] : src.data[i];
        lo = std::min(lo, v); hi = std::max(hi, v);
    }
    if (mn) *mn = lo;
    if (mx) *mx = hi;
}
void split(const Mat &src, std::vector<Mat> &mv) {
    const int cn = src.channels();
    mv.resize(cn);
    for (int c = 0; c < cn; c++) {
        mv[c].create(src.rows, src.cols, CV_MAKETYPE(src.depth(), 1));
        for (size_t i = 0; i < (size_t)src.rows * src.cols; i++)
            reinterpret_cast<float *>(mv[c].data)[i] = reinterpret_cast<const float *>(src.data)[i * cn + c];
    }
}
void merge(const std::vector<Mat> &mv, Mat &dst) {
    const int cn = (int)mv.size();
    Mat out(mv[0].rows, mv[0].cols, CV_MAKETYPE(mv[0].depth(), cn));
    for (int c = 0; c < cn; c++)
        for (size_t i = 0; i < (size_t)out.rows * out.cols; i++)
            reinterpret_cast<float *>(out.data)[i * cn + c] = reinterpret_cast<const float *>(mv[c].data)[i];
    dst = out;
}
void namedWindow(const String &, int) {}
void moveWindow(const String &, int, int) {}
void resizeWindow(const String &, int, int) {}
void imshow(const String &, const Mat &) {}
int waitKey(int) { return 0; }
bool imwrite(const String &, const Mat &, const std::vector<int> &) { return true; }
Mat imread(const String &, int) { return Mat(); }

} // namespace cv
