// match.cu -- greedy score-ordered matching over a distance matrix (second half).
//
// Replaces the pair walk of ScoreMatcher.match + BaseMatcher.match_by_order (reference d3d/tracking/matcher.pyx:93-122, 138-162), which
// the detection evaluator runs once per score threshold on the distance cache of prepare_boxes (d3d/benchmarks.pyx:220-238): source
// boxes are visited from the best score down, each takes the closest destination box of its own category that is still free and not
// farther than the category's threshold.  The walk is sequential in the sources, so one CTA owns one threshold set (the evaluator's
// 40 thresholds are 40 CTAs of one launch) and spends two barriers per source: all threads scan the row for the nearest free
// candidate, one thread commits it.  Ties in the distance go to the lower destination index (the reference's np.argsort leaves them
// unspecified).
#include "common.cuh"

namespace d3d {

constexpr int MT_THREADS = 256;

__device__ __forceinline__ unsigned long long mt_key(float d, uint32_t j)
{
    uint32_t u = __float_as_uint(d);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending order-preserving
    return ((unsigned long long)u << 32) | j;
}

__global__ void __launch_bounds__(MT_THREADS) match_greedy_kernel(const float *__restrict__ dist, int64_t n, int64_t m, int64_t ld, const int32_t *__restrict__ src_order,
                                                                  const int32_t *__restrict__ src_tag, const int32_t *__restrict__ dst_tag,
                                                                  const float *__restrict__ thresholds, int32_t ncat, int32_t *__restrict__ src_assign,
                                                                  int32_t *__restrict__ dst_assign)
{
    extern __shared__ unsigned char taken[];   // [m] destination already assigned
    __shared__ unsigned long long s_best[MT_THREADS / 32];
    const int t = blockIdx.x, tid = threadIdx.x;
    int32_t *sa = src_assign + (int64_t)t * n, *da = dst_assign + (int64_t)t * m;
    for (int64_t j = tid; j < m; j += MT_THREADS) { taken[j] = 0; da[j] = -1; }
    for (int64_t i = tid; i < n; i += MT_THREADS) sa[i] = -1;
    __syncthreads();
    for (int64_t p = 0; p < n; p++) {
        const int32_t i = src_order[p];
        const int32_t tag = src_tag[i];
        unsigned long long best = ~0ull;
        if (tag >= 0 && tag < ncat) {
            const float thr = thresholds[(int64_t)t * ncat + tag];
            const float *row = dist + (int64_t)i * ld;
            for (int64_t j = tid; j < m; j += MT_THREADS) {
                if (!taken[j] && dst_tag[j] == tag) {
                    const float d = row[j];
                    if (d <= thr) { const unsigned long long k = mt_key(d, (uint32_t)j); best = k < best ? k : best; }   // NaN never matches, like the reference's <=
                }
            }
        }
#pragma unroll
        for (int s = 16; s; s >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, s); best = o < best ? o : best; }
        if ((tid & 31) == 0) s_best[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            unsigned long long b = s_best[0];
#pragma unroll
            for (int w = 1; w < MT_THREADS / 32; w++) b = s_best[w] < b ? s_best[w] : b;
            if (b != ~0ull) { const uint32_t j = (uint32_t)b; taken[j] = 1; sa[i] = (int32_t)j; da[j] = i; }
        }
        __syncthreads();
    }
}

}  // namespace d3d

using namespace d3d;
extern "C" int d3d_match_greedy_f32(const float *dist, int64_t n, int64_t m, int64_t ld, const int32_t *src_order, const int32_t *src_tag, const int32_t *dst_tag,
                                    const float *thresholds, int32_t nsets, int32_t ncat, int32_t *src_assign, int32_t *dst_assign, void *stream)
{
    if (n < 0 || m < 0 || nsets < 0 || ncat < 1 || ld < m) return D3D_ERR_INVALID_ARGUMENT;
    if (nsets == 0) return D3D_OK;
    if (!thresholds || (n > 0 && (!src_order || !src_tag || !src_assign)) || (m > 0 && (!dst_tag || !dst_assign)) || (n > 0 && m > 0 && !dist)) return D3D_ERR_INVALID_ARGUMENT;
    if (m > 200 * 1024) return D3D_ERR_UNSUPPORTED;   // one flag byte per destination box in shared memory
    const size_t smem = (size_t)(m > 0 ? m : 1);
    if (smem > 40 * 1024) D3D_CUDA_TRY(cudaFuncSetAttribute(match_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_greedy_kernel<<<(unsigned)nsets, MT_THREADS, smem, (cudaStream_t)stream>>>(dist, n, m, ld, src_order, src_tag, dst_tag, thresholds, ncat, src_assign, dst_assign);
    D3D_LAUNCHED();
    return D3D_OK;
}
