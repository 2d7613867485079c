// crop.cu -- point-in-rotated-box mask.
//
// Replaces reference crop_2dr (d3d/box/utils.h:45, utils.cpp:10-47; the reference has no CUDA version of it): for M boxes
// and N points, mask[i][j] = the AABB of box i strictly contains point j AND no edge of the quad has the point on its
// right (dgal Poly2::contains, thirdparty/dgal/geometry.hpp:218-229 over _cross :160-163; vertices as
// poly2_from_xywhr :417-429, bounding box as aabox2_from_poly2 :398-414 with the open test of :185-188).
//
// Layout / roofline: the output is bool[M, N] row-major = 1 byte per pair; a thread owns 16 consecutive points and
// writes their 16 mask bytes of one box as a single 16-byte store, a warp writes 512 contiguous bytes of a mask row.
// The box tile's 12 numbers per box (4 vertices + AABB) sit in shared memory.  ~35 flops per pair against 1 stored
// byte: FP32 issue bound for float (59 TFLOP/s -> ~1.7 T pairs/s), nearer the HBM store stream for small M.
// The per-pair arithmetic uses the explicit round-to-nearest intrinsics: the reference is compiled without FMA
// contraction and the mask is compared bit for bit.
#include "common.cuh"

namespace d3d {

template <typename T> struct CropBox { T vx[4], vy[4], minx, maxx, miny, maxy; };

template <typename T> __device__ __forceinline__ T rn_mul(T a, T b);
template <> __device__ __forceinline__ float rn_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double rn_mul<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T rn_sub(T a, T b);
template <> __device__ __forceinline__ float rn_sub<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double rn_sub<double>(double a, double b) { return __dsub_rn(a, b); }
template <typename T> __device__ __forceinline__ T rn_add(T a, T b);
template <> __device__ __forceinline__ float rn_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double rn_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T rn_div2(T a);
template <> __device__ __forceinline__ float rn_div2<float>(float a) { return __fmul_rn(a, 0.5f); }     // x / 2 is exact
template <> __device__ __forceinline__ double rn_div2<double>(double a) { return __dmul_rn(a, 0.5); }

template <typename T>
__global__ void __launch_bounds__(256) crop_prep_kernel(const T *__restrict__ boxes, int64_t m, CropBox<T> *__restrict__ recs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const T x = boxes[5 * i], y = boxes[5 * i + 1], w = boxes[5 * i + 2], h = boxes[5 * i + 3], r = boxes[5 * i + 4];
    T sn, cs;
    if (sizeof(T) == 4) { sn = (T)sinf((float)r); cs = (T)cosf((float)r); } else { sn = (T)sin((double)r); cs = (T)cos((double)r); }
    const T dxsin = rn_div2(rn_mul(w, sn)), dxcos = rn_div2(rn_mul(w, cs)), dysin = rn_div2(rn_mul(h, sn)), dycos = rn_div2(rn_mul(h, cs));
    CropBox<T> b;
    b.vx[0] = rn_add(rn_sub(x, dxcos), dysin); b.vy[0] = rn_sub(rn_sub(y, dxsin), dycos);
    b.vx[1] = rn_add(rn_add(x, dxcos), dysin); b.vy[1] = rn_sub(rn_add(y, dxsin), dycos);
    b.vx[2] = rn_sub(rn_add(x, dxcos), dysin); b.vy[2] = rn_add(rn_add(y, dxsin), dycos);
    b.vx[3] = rn_sub(rn_sub(x, dxcos), dysin); b.vy[3] = rn_add(rn_sub(y, dxsin), dycos);
    b.minx = b.maxx = b.vx[0]; b.miny = b.maxy = b.vy[0];
#pragma unroll
    for (int k = 1; k < 4; k++) {
        b.minx = b.vx[k] < b.minx ? b.vx[k] : b.minx; b.maxx = b.vx[k] > b.maxx ? b.vx[k] : b.maxx;
        b.miny = b.vy[k] < b.miny ? b.vy[k] : b.miny; b.maxy = b.vy[k] > b.maxy ? b.vy[k] : b.maxy;
    }
    recs[i] = b;
}

constexpr int CROP_THREADS = 256, CROP_PPT = 16, CROP_BT = 16;   // points per thread, boxes per CTA

template <typename T>
__global__ void __launch_bounds__(CROP_THREADS) crop2dr_kernel(const T *__restrict__ pts, int64_t n, const CropBox<T> *__restrict__ recs, int64_t m,
                                                               uint8_t *__restrict__ mask)
{
    __shared__ CropBox<T> sb[CROP_BT];
    const int64_t box0 = (int64_t)blockIdx.y * CROP_BT;
    const int nb = (int)(m - box0 < CROP_BT ? m - box0 : CROP_BT);
    for (int i = threadIdx.x; i < nb * (int)(sizeof(CropBox<T>) / sizeof(T)); i += CROP_THREADS)
        reinterpret_cast<T *>(sb)[i] = reinterpret_cast<const T *>(recs + box0)[i];
    __syncthreads();
    const int64_t p0 = ((int64_t)blockIdx.x * CROP_THREADS + threadIdx.x) * CROP_PPT;
    if (p0 >= n) return;
    T px[CROP_PPT], py[CROP_PPT];
    const bool full = p0 + CROP_PPT <= n;
#pragma unroll
    for (int k = 0; k < CROP_PPT; k++) {
        const int64_t j = p0 + k < n ? p0 + k : n - 1;
        px[k] = pts[2 * j]; py[k] = pts[2 * j + 1];
    }
    const bool vec = full && (n % 16 == 0) && ((reinterpret_cast<uintptr_t>(mask) & 15) == 0);
    for (int b = 0; b < nb; b++) {
        const CropBox<T> B = sb[b];
        uint32_t wds[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < CROP_PPT; k++) {
            bool in = px[k] > B.minx && px[k] < B.maxx && py[k] > B.miny && py[k] < B.maxy;
            if (in) {
#pragma unroll
                for (int e = 0; e < 4; e++) {   // edge a -> e in the reference's order 3->0, 0->1, 1->2, 2->3
                    const int a = (e + 3) & 3;
                    const T c = rn_sub(rn_mul(rn_sub(B.vx[e], B.vx[a]), rn_sub(py[k], B.vy[e])), rn_mul(rn_sub(B.vy[e], B.vy[a]), rn_sub(px[k], B.vx[e])));
                    in = in && !(c < T(0));
                }
            }
            wds[k >> 2] |= (in ? 1u : 0u) << ((k & 3) * 8);
        }
        uint8_t *row = mask + (box0 + b) * n + p0;
        if (vec) {
            __stcs(reinterpret_cast<uint4 *>(row), make_uint4(wds[0], wds[1], wds[2], wds[3]));
        } else {
            for (int k = 0; k < CROP_PPT && p0 + k < n; k++) row[k] = (uint8_t)((wds[k >> 2] >> ((k & 3) * 8)) & 1u);
        }
    }
}

template <typename T> static size_t crop_ws_bytes(int64_t m) { return align_up((size_t)(m > 0 ? m : 1) * sizeof(CropBox<T>)) + 256; }

template <typename T>
static int crop_impl(const T *pts, int64_t n, const T *boxes, int64_t m, uint8_t *mask, void *ws, size_t ws_bytes, cudaStream_t st)
{
    if (n < 0 || m < 0) return D3D_ERR_INVALID_ARGUMENT;
    if (n == 0 || m == 0) return D3D_OK;
    if (!pts || !boxes || !mask) return D3D_ERR_INVALID_ARGUMENT;
    if (!ws || ws_bytes < crop_ws_bytes<T>(m)) return D3D_ERR_WORKSPACE;
    const int64_t gy = cdiv(m, CROP_BT), gx = cdiv(n, (int64_t)CROP_THREADS * CROP_PPT);
    if (gy > 65535 || gx > 0x7fffffffll) return D3D_ERR_INVALID_ARGUMENT;   // up to ~1M boxes per call
    CropBox<T> *recs = reinterpret_cast<CropBox<T> *>(ws);
    crop_prep_kernel<T><<<(unsigned)cdiv(m, 256), 256, 0, st>>>(boxes, m, recs); D3D_LAUNCHED();
    crop2dr_kernel<T><<<dim3((unsigned)gx, (unsigned)gy), CROP_THREADS, 0, st>>>(pts, n, recs, m, mask); D3D_LAUNCHED();
    return D3D_OK;
}

}  // namespace d3d

using namespace d3d;
extern "C" size_t d3d_crop2dr_workspace_bytes(int64_t m, int dtype) { return dtype == D3D_F64 ? crop_ws_bytes<double>(m) : crop_ws_bytes<float>(m); }
extern "C" int d3d_crop2dr_f32(const float *points, int64_t n, const float *boxes, int64_t m, uint8_t *mask, void *ws, size_t wsb, void *stream)
{ return crop_impl<float>(points, n, boxes, m, mask, ws, wsb, (cudaStream_t)stream); }
extern "C" int d3d_crop2dr_f64(const double *points, int64_t n, const double *boxes, int64_t m, uint8_t *mask, void *ws, size_t wsb, void *stream)
{ return crop_impl<double>(points, n, boxes, m, mask, ws, wsb, (cudaStream_t)stream); }
