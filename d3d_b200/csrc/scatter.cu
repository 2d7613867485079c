// scatter.cu -- aligned_scatter forward (gather at fractional coordinates) and backward (scatter-add).
//
// Replaces reference aligned_scatter_forward/backward[_cuda] (d3d/point/scatter.cpp:34-201,
// d3d/point/scatter_cuda.cu:37-241).  The reference launches one 512-thread block per point with at
// most `nchan` threads active and rebuilds the neighbour table in shared memory with atomMul; here one
// thread owns one (point, channel) output with the channel index fastest, so the N x C output is written
// fully coalesced, the 2^dim neighbour offsets/weights live in registers, and the grid is sized by work,
// not by point count.  Forward arithmetic keeps the reference CPU order (sequential sum over the
// neighbours, separate multiply and add) so the fp32/fp64 results are bit-identical to it.
#include "common.cuh"

namespace d3d {

struct ScDims { int d[3]; long long plane; };

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// neighbour j (bit d of j = take the upper neighbour along dim d): linear offset in the plane and weight.
// Mirrors _fill_lcoords (scatter.cpp:34-77) including its quirk that an integral in-range coordinate
// gives both coincident neighbours weight 1.
template <typename T, int DIM, bool LINEAR>
__device__ __forceinline__ void neighbours(const T *__restrict__ crd, const ScDims &dm, long long off[1 << DIM], T wgt[1 << DIM])
{
    int lo[DIM], hi[DIM]; T wlo[DIM], whi[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        const T x = crd[d + 1];
        const int dmax = dm.d[d] - 1;
        if (x > (T)dmax) { lo[d] = hi[d] = dmax; wlo[d] = whi[d] = T(0.5); }
        else if (x < (T)0) { lo[d] = hi[d] = 0; wlo[d] = whi[d] = T(0.5); }
        else {
            int k = (int)x;            // truncation; x >= 0 here so this is floor except for the checks below
            int fl = k > x ? k - 1 : k;   // _floor, scatter.cpp:22-27
            int ce = k < x ? k + 1 : k;   // _ceil,  scatter.cpp:28-33
            lo[d] = fl; hi[d] = ce;
            wlo[d] = add_rn<T>(add_rn<T>(T(1), -x), (T)fl);   // 1 - x + floor
            whi[d] = add_rn<T>(add_rn<T>(T(1), x), -(T)ce);   // 1 + x - ceil
        }
    }
#pragma unroll
    for (int j = 0; j < (1 << DIM); j++) {
        long long o = 0; T w = T(1);
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const bool up = (j >> d) & 1;
            o = o * dm.d[d] + (up ? hi[d] : lo[d]);
            if (LINEAR) w = mul_rn<T>(w, up ? whi[d] : wlo[d]);
        }
        off[j] = o; wgt[j] = w;
    }
}

template <typename T, int DIM, bool LINEAR>
__global__ void __launch_bounds__(256) scatter_fwd_kernel(const T *__restrict__ coord, int64_t n, const T *__restrict__ image, int64_t nchan, ScDims dm, T *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nchan) return;
    const int64_t i = t / nchan, c = t % nchan;
    const T *crd = coord + i * (DIM + 1);
    long long off[1 << DIM]; T wgt[1 << DIM];
    neighbours<T, DIM, LINEAR>(crd, dm, off, wgt);
    const int b = (int)crd[0];
    const T *pl = image + ((int64_t)b * nchan + c) * dm.plane;
    T v[1 << DIM];
#pragma unroll
    for (int j = 0; j < (1 << DIM); j++) v[j] = __ldg(pl + off[j]);   // all gathers in flight before the sum
    T sum = T(0);
#pragma unroll
    for (int j = 0; j < (1 << DIM); j++) sum = add_rn<T>(sum, LINEAR ? mul_rn<T>(v[j], wgt[j]) : v[j]);
    out[t] = LINEAR ? sum : sum / (T)(1 << DIM);
}

template <typename T, int DIM, bool LINEAR>
__global__ void __launch_bounds__(256) scatter_bwd_kernel(const T *__restrict__ coord, int64_t n, const T *__restrict__ grad, int64_t nchan, ScDims dm, T *__restrict__ image_grad)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nchan) return;
    const int64_t i = t / nchan, c = t % nchan;
    const T *crd = coord + i * (DIM + 1);
    long long off[1 << DIM]; T wgt[1 << DIM];
    neighbours<T, DIM, LINEAR>(crd, dm, off, wgt);
    const int b = (int)crd[0];
    T *pl = image_grad + ((int64_t)b * nchan + c) * dm.plane;
    const T g = grad[t];
#pragma unroll
    for (int j = 0; j < (1 << DIM); j++)
        atomicAdd(pl + off[j], LINEAR ? mul_rn<T>(g, wgt[j]) : g / (T)(1 << DIM));   // RED.ADD, no return value
}

template <typename T, int DIM>
static int scatter_launch(bool fwd, const T *coord, int64_t n, const T *src, int64_t nchan, ScDims dm, int align, T *dst, cudaStream_t st)
{
    const int64_t work = n * nchan;
    if (work == 0) return D3D_OK;
    const unsigned blocks = (unsigned)cdiv(work, 256);
    if (fwd) {
        if (align == D3D_ALIGN_LINEAR) scatter_fwd_kernel<T, DIM, true><<<blocks, 256, 0, st>>>(coord, n, src, nchan, dm, dst);
        else scatter_fwd_kernel<T, DIM, false><<<blocks, 256, 0, st>>>(coord, n, src, nchan, dm, dst);
    } else {
        if (align == D3D_ALIGN_LINEAR) scatter_bwd_kernel<T, DIM, true><<<blocks, 256, 0, st>>>(coord, n, src, nchan, dm, dst);
        else scatter_bwd_kernel<T, DIM, false><<<blocks, 256, 0, st>>>(coord, n, src, nchan, dm, dst);
    }
    D3D_LAUNCHED();
    return D3D_OK;
}

template <typename T>
static int scatter_impl(bool fwd, const void *coord, int64_t n, int32_t dim, const void *src, int64_t nbatch, int64_t nchan, const int64_t *dims, int align, void *dst,
                        cudaStream_t st)
{
    if (align != D3D_ALIGN_MEAN && align != D3D_ALIGN_LINEAR) return D3D_ERR_INVALID_ARGUMENT;   // reference: "Unsupported align type!"
    if (dim < 1 || dim > 3) return D3D_ERR_INVALID_ARGUMENT;                                    // reference: "Unsupported dimension size"
    if (n < 0 || nbatch < 0 || nchan < 0 || !dims) return D3D_ERR_INVALID_ARGUMENT;
    if (n * nchan > 0 && (!coord || !src || !dst)) return D3D_ERR_INVALID_ARGUMENT;
    ScDims dm; dm.plane = 1;
    for (int d = 0; d < 3; d++) { dm.d[d] = d < dim ? (int)dims[d] : 1; if (d < dim) { if (dims[d] <= 0 || dims[d] > 0x7fffffff) return D3D_ERR_INVALID_ARGUMENT; dm.plane *= dims[d]; } }
    if (n * nchan >= (1ll << 31) * 256) return D3D_ERR_INVALID_ARGUMENT;
    switch (dim) {
    case 1: return scatter_launch<T, 1>(fwd, (const T *)coord, n, (const T *)src, nchan, dm, align, (T *)dst, st);
    case 2: return scatter_launch<T, 2>(fwd, (const T *)coord, n, (const T *)src, nchan, dm, align, (T *)dst, st);
    default: return scatter_launch<T, 3>(fwd, (const T *)coord, n, (const T *)src, nchan, dm, align, (T *)dst, st);
    }
}

}  // namespace d3d

using namespace d3d;
extern "C" int d3d_aligned_scatter_forward(const void *coord, int64_t n, int32_t dim, const void *image, int64_t nbatch, int64_t nchan, const int64_t *dims_host, int align,
                                           int dtype, void *out, void *stream)
{
    return dtype == D3D_F64 ? scatter_impl<double>(true, coord, n, dim, image, nbatch, nchan, dims_host, align, out, (cudaStream_t)stream)
                            : scatter_impl<float>(true, coord, n, dim, image, nbatch, nchan, dims_host, align, out, (cudaStream_t)stream);
}
extern "C" int d3d_aligned_scatter_backward(const void *coord, int64_t n, int32_t dim, const void *grad, int64_t nbatch, int64_t nchan, const int64_t *dims_host, int align,
                                            int dtype, void *image_grad, void *stream)
{
    return dtype == D3D_F64 ? scatter_impl<double>(false, coord, n, dim, grad, nbatch, nchan, dims_host, align, image_grad, (cudaStream_t)stream)
                            : scatter_impl<float>(false, coord, n, dim, grad, nbatch, nchan, dims_host, align, image_grad, (cudaStream_t)stream);
}
