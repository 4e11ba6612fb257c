#pragma once
#include "core.hpp"
namespace cv {
// restated in oracle/ref_glue/shim_impl.cpp, pinned to python cv2 4.13
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REPLICATE);
void resize(const Mat &src, Mat &dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
} // namespace cv
