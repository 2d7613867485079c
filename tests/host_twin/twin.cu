// twin.cu -- TEST INFRASTRUCTURE: compiles the product's clip math (d3d_b200/csrc/geom.cuh) as HOST code
// so the algorithm can be checked against the CPU oracle on a machine without a GPU.  Nothing in the
// product loads this library.
#include "../../d3d_b200/csrc/geom.cuh"
using namespace d3d;
template <typename T> static void run(const T *b1, long n, const T *b2, long m, T *out, int aabb)
{
    for (long i = 0; i < n; i++)
        for (long j = 0; j < m; j++) {
            const T *p = b1 + 5 * i, *q = b2 + 5 * j;
            if (aabb) out[i * m + j] = aabb_iou(make_aabb_rec<T>(p[0], p[1], p[2], p[3], p[4]), make_aabb_rec<T>(q[0], q[1], q[2], q[3], q[4]));
            else {
                BoxRec<T> A = make_box_rec<T>(p[0], p[1], p[2], p[3], p[4]), B = make_box_rec<T>(q[0], q[1], q[2], q[3], q[4]);
                T ddx = A.cx - B.cx, ddy = A.cy - B.cy, rs = A.rho + B.rho;
                out[i * m + j] = (ddx * ddx + ddy * ddy <= rs * rs) ? rbox_iou<T>(A, B) : T(0);
            }
        }
}
extern "C" void twin_iou_f32(const float *b1, long n, const float *b2, long m, float *out, int aabb) { run<float>(b1, n, b2, m, out, aabb); }
extern "C" void twin_iou_f64(const double *b1, long n, const double *b2, long m, double *out, int aabb) { run<double>(b1, n, b2, m, out, aabb); }
