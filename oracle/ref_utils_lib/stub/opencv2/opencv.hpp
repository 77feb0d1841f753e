// TEST INFRASTRUCTURE ONLY -- the few OpenCV names upsample_velodyne (utils_lib.cpp:20-84) mentions, enough
// for the translation unit to compile.  The eval path (params["upsample"] = 0, kittiloader/kitti.py:693-697)
// never reaches them; the shim refuses upsample != 0.
#pragma once
#include <vector>
#define CV_32FC1 5
namespace cv {
struct Scalar { double v; Scalar(double x = 0) : v(x) {} };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
enum { INTER_NEAREST = 0 };
struct Mat {
    int rows, cols;
    std::vector<float> d;
    Mat() : rows(0), cols(0) {}
    Mat(int r, int c, int, Scalar s) : rows(r), cols(c), d((size_t)r * c, (float)s.v) {}
    template <class T> T& at(int i, int j) { return d[(size_t)i * cols + j]; }
    Size size() const { return Size(cols, rows); }
};
inline void resize(const Mat& src, Mat& dst, Size, double fx, double fy, int) {   // nearest neighbour
    Mat o((int)(src.rows * fy + 0.5), (int)(src.cols * fx + 0.5), CV_32FC1, Scalar(0.));
    for (int i = 0; i < o.rows; ++i)
        for (int j = 0; j < o.cols; ++j) {
            int si = (int)(i / fy), sj = (int)(j / fx);
            if (si >= src.rows) si = src.rows - 1;
            if (sj >= src.cols) sj = src.cols - 1;
            o.d[(size_t)i * o.cols + j] = src.d[(size_t)si * src.cols + sj];
        }
    dst = o;
}
}  // namespace cv
