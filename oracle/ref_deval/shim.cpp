// C entry points around the reference's own depthError / evaluateErrors
// (external/deval_lib/src/evaluate_depth.h:19-146, compiled from /root/reference, see Makefile).
// This replaces the reference's pybind module (external/deval_lib/python/pyevaluatedepth_lib.cpp),
// which needs OpenCV / Eigen / pybind11 headers that are not in this image.  TEST INFRASTRUCTURE ONLY.
#include <evaluate_depth.h>

extern "C" int ref_depth_error(const float* first, const float* second, int width, int height, float* out9) {
    // same construction as the pybind type_caster: DepthImage(dataPtr, shape[1], shape[0])
    DepthImage a(first, width, height), b(second, width, height);
    try {
        std::vector<float> e = depthError(a, b);
        for (int i = 0; i < 9; ++i) out9[i] = e[i];
        return 0;
    } catch (int) {
        return 1;
    }
}

// errs [n][9] -> out [9][3] = (mean, min, max) per metric, in the order of evaluate_depth.h:123-131
extern "C" void ref_evaluate_errors(const float* errs, int n, float* out27) {
    static const char* metrics[] = {"mae", "rmse", "inverse mae", "inverse rmse", "log mae", "log rmse",
                                    "scale invariant log", "abs relative", "squared relative"};
    std::vector<std::vector<float> > v(n, std::vector<float>(9));
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 9; ++j) v[i][j] = errs[i * 9 + j];
    std::map<std::string, std::vector<float> > r = evaluateErrors(v);
    for (int j = 0; j < 9; ++j)
        for (int s = 0; s < 3; ++s) out27[j * 3 + s] = r[metrics[j]][s];
}
