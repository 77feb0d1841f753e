/* CPU restatement of the reference's LiDAR depth-map generation -- TEST INFRASTRUCTURE ONLY.
 *
 *   generate_depth   external/utils_lib/python/utils_lib.cpp:86-160 (upsample = 0 branch: the eval
 *                    default, kittiloader/kitti.py:693-697) : velo -> camera frame (:95), drop z < 0.1
 *                    (:98-107), project (:115-118), z-buffer (:121-130), neighbourhood filter (:133-157)
 *   minpool          utils/img_utils.py:87-95 as called at kittiloader/kitti.py:706 (scale 4, default 1000)
 *
 * Pinning of generate_depth: the reference ships no test or golden vector for it, and its build needs Eigen,
 * OpenCV and pybind11, none of which is in this image.  oracle/ref_utils_lib/ therefore compiles the
 * reference's own source file, from where it lies, against minimal stand-ins for those three libraries
 * (oracle/_ref/libutils_ref.so); this restatement agrees with that build bit for bit on the committed cases
 * and on random clouds (tests/test_lidar.py), which pins its control flow, comparisons, int casts, z-buffer
 * and filter.  What stays a CHOICE is the association of the 4-term sums in the two small matrix products,
 * which Eigen leaves to its kernels: row times column, left to right, separate multiply and add (build with
 * -ffp-contract=off) -- the stand-in makes the same choice, so this part is stated, not pinned.  minpool IS
 * pinned: the reference's Python runs in the build container (tests/golden/make_lidar_golden.py).
 */
#include <stdlib.h>
#include <string.h>

static float dot4(const float* m, const float* p) {
    float s = m[0] * p[0];
    s = s + m[1] * p[1];
    s = s + m[2] * p[2];
    s = s + m[3] * p[3];
    return s;
}

/* velo [n][4] (x, y, z, 1), intr [3][4], m_velo2cam [4][4], all row-major.  out [height][width]. */
void oracle_generate_depth(const float* velo, int n, const float* intr, const float* m_velo2cam, int width,
                           int height, int filtering, float filterdiff, float* out) {
    float* raw = (float*)calloc((size_t)width * height, sizeof(float));
    for (int i = 0; i < n; ++i) {
        float cam[4];
        for (int r = 0; r < 4; ++r) cam[r] = dot4(m_velo2cam + 4 * r, velo + 4 * i);      /* :95 */
        if (!(cam[2] >= 0.1f)) continue;                                                    /* :101 */
        float proj[3];
        for (int r = 0; r < 3; ++r) proj[r] = dot4(intr + 4 * r, cam);                      /* :115 */
        const float px = proj[0] / proj[2], py = proj[1] / proj[2];                         /* :116-117 */
        const int u = (int)(px - 0.5), v = (int)(py - 0.5);                                 /* :123-124, double */
        if (u < 0 || u >= width || v < 0 || v >= height) continue;
        const float z = cam[2], cur = raw[v * width + u];                                   /* :118, :126-129 */
        if (z < cur || cur == 0) raw[v * width + u] = z;
    }
    memset(out, 0, sizeof(float) * (size_t)width * height);
    const int off = filtering;
    for (int v = off; v < height - off - 1; ++v)                                            /* :136-137 */
        for (int u = off; u < width - off - 1; ++u) {
            const float z = raw[v * width + u];
            int bad = 0;
            for (int vv = v - off; vv < v + off + 1; ++vv)
                for (int uu = u - off; uu < u + off + 1; ++uu) {
                    if (vv == v && uu == u) continue;
                    const float zn = raw[vv * width + uu];
                    if (zn == 0) continue;
                    if ((zn - z) < -filterdiff) bad = 1;                                    /* :148-151 */
                }
            if (!bad) out[v * width + u] = z;
        }
    free(raw);
}

/* -max_pool2d(-x, scale) with zeros replaced by `dflt` first and `dflt` mapped back to 0 afterwards
 * (utils/img_utils.py:88-92); floor division of the size, as max_pool2d does. */
void oracle_minpool(const float* in, int width, int height, int scale, float dflt, float* out) {
    const int w2 = width / scale, h2 = height / scale;
    for (int y = 0; y < h2; ++y)
        for (int x = 0; x < w2; ++x) {
            float m = 0.f;
            int first = 1;
            for (int dy = 0; dy < scale; ++dy)
                for (int dx = 0; dx < scale; ++dx) {
                    float v = in[(y * scale + dy) * width + x * scale + dx];
                    if (v == 0) v = dflt;
                    if (first || v < m) { m = v; first = 0; }
                }
            out[y * w2 + x] = (m == dflt) ? 0.f : m;
        }
}
