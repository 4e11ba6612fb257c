#pragma once
#include "core.hpp"
namespace cv {
// display / file output: no-ops (only reached under verbose bits the oracle never sets)
void namedWindow(const String &name, int flags = 0);
void moveWindow(const String &name, int x, int y);
void resizeWindow(const String &name, int w, int h);
void imshow(const String &name, const Mat &m);
int waitKey(int delay = 0);
bool imwrite(const String &file, const Mat &m, const std::vector<int> &params = std::vector<int>());
Mat imread(const String &file, int flags = 1);
} // namespace cv
