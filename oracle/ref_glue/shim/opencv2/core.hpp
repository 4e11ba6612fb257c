// OpenCV-lite: just enough of the cv:: API surface for the reference's variational_mt.cpp,
// utils/utils.h and utils/parameter_list.{h,cpp} to compile UNMODIFIED.  TEST INFRASTRUCTURE ONLY
// (oracle/_ref build).  OpenCV itself is not available offline; the two functions with real
// arithmetic on the hot path (GaussianBlur, resize) are restated in ../shim_impl.cpp and pinned to
// python cv2 4.13 by tests/test_oracle_pin.py.
#pragma once
#include <stdint.h>
#include <string.h>
#include <cmath>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_IMWRITE_PXM_BINARY 32

namespace cv {

typedef std::string String;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> std::ostream &operator<<(std::ostream &os, const Point_<T> &p) { return os << "[" << p.x << ", " << p.y << "]"; }

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T, int N> struct Vec {
    T val[N];
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<uchar, 3> Vec3b;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { WINDOW_NORMAL = 0, WINDOW_FREERATIO = 0x100 };

class Mat {
public:
    int rows, cols;
    uchar *data;
    Mat() : rows(0), cols(0), data(nullptr), type_(0) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * elemSize(), 0);
        data = buf_->data();
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_32F ? 4 : 1); }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    Size size() const { return Size(cols, rows); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        if (data) memcpy(m.data, data, (size_t)rows * cols * elemSize());
        return m;
    }
    template <typename T> T &at(int i, int j) { return reinterpret_cast<T *>(data)[(size_t)i * cols + j]; }
    template <typename T> const T &at(int i, int j) const { return reinterpret_cast<const T *>(data)[(size_t)i * cols + j]; }
    void convertTo(Mat &dst, int rtype, double alpha = 1, double beta = 0) const;

private:
    int type_;
    std::shared_ptr<std::vector<uchar>> buf_;
};

Mat operator+(const Mat &a, double s);
Mat operator-(const Mat &a, double s);
Mat operator*(const Mat &a, double s);
Mat operator*(double s, const Mat &a);
Mat operator/(const Mat &a, double s);

void minMaxLoc(const Mat &src, double *minVal, double *maxVal = nullptr, Point *minLoc = nullptr, Point *maxLoc = nullptr);
void split(const Mat &src, std::vector<Mat> &mv);
void merge(const std::vector<Mat> &mv, Mat &dst);

} // namespace cv

typedef cv::Scalar CvScalar;
