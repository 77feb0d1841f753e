// C entry point around the reference's own generate_depth (external/utils_lib/python/utils_lib.cpp:86-160,
// compiled from /root/reference, see Makefile) -- TEST INFRASTRUCTURE ONLY.  Eigen, OpenCV and pybind11 are not
// in this image; stub/ holds minimal stand-ins (see stub/Eigen/Dense for the one choice they make).
#include <utils_lib.cpp>

// velo [n][4], intr [3][4], m_velo2cam [4][4], out [height][width], all row-major float32.
extern "C" int ref_generate_depth(const float* velo, int n, const float* intr, const float* m_velo2cam, int width,
                                  int height, int filtering, float filterdiff, float* out) {
    Eigen::MatrixXf V(n, 4), I(3, 4), M(4, 4);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 4; ++j) V(i, j) = velo[i * 4 + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) I(i, j) = intr[i * 4 + j];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) M(i, j) = m_velo2cam[i * 4 + j];
    py::dict params;
    params["upsample"].v = 0;            // the eval branch (kittiloader/kitti.py:693-697)
    params["filtering"].v = filtering;
    params["filterdiff"].v = filterdiff;
    const Eigen::MatrixXf d = generate_depth(V, I, M, width, height, params);
    if (d.rows() != height || d.cols() != width) return 1;
    for (int v = 0; v < height; ++v)
        for (int u = 0; u < width; ++u) out[v * width + u] = d(v, u);
    return 0;
}
