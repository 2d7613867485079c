// prims.cu -- device-wide primitives written for this library: exclusive scan and a stable LSD radix
// sort of (key, value) pairs.  They serve NMS (score ordering) and voxelization (grouping points by
// voxel key deterministically).  No CUB/Thrust: everything launched here is our own kernel.
#include "prims.cuh"
#include <cooperative_groups.h>

namespace d3d {

// ------------------------------------------------------------------------------------------------
// exclusive scan of u32 -> u32 (sums must fit 32 bits: callers scan flags / small counts)
// three phases: per-tile reduce, single-block scan of tile sums, per-tile scan + add
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (unsigned)d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (blockDim = SCAN_THREADS); returns exclusive prefix, total in *tot
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *tot)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t total;
    uint32_t inc = warp_incl_scan(v);
    unsigned w = threadIdx.x >> 5;
    if (lane_id() == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane_id() < SCAN_THREADS / 32 ? wsum[lane_id()] : 0;
        uint32_t si = warp_incl_scan(s);
        if (lane_id() < SCAN_THREADS / 32) wsum[lane_id()] = si - s;
        if (lane_id() == SCAN_THREADS / 32 - 1) total = si;
    }
    __syncthreads();
    uint32_t r = inc - v + wsum[w];
    *tot = total;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ tile_sums)
{
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// one block scans all tile sums in place (exclusive); writes the grand total to total_out (may be null)
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(uint32_t *__restrict__ tile_sums, int64_t ntiles, uint32_t *__restrict__ total_out)
{
    uint32_t carry = 0;
    for (int64_t base = 0; base < ntiles; base += SCAN_THREADS) {
        int64_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0, tot;
        uint32_t ex = block_excl_scan(v, &tot);
        if (i < ntiles) tile_sums[i] = ex + carry;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t *in, int64_t n, const uint32_t *__restrict__ tile_sums, uint32_t *out)
{
    // thread owns SCAN_ITEMS consecutive elements (blocked arrangement) so the scan order is index order
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, &tot) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}

size_t scan_workspace_bytes(int64_t n) { return align_up((size_t)(cdiv(n, SCAN_TILE) + 1) * sizeof(uint32_t)); }

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total_out, void *ws, cudaStream_t st)
{
    if (n <= 0) {
        if (total_out) D3D_CUDA_TRY(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), st));
        return D3D_OK;
    }
    int64_t ntiles = cdiv(n, SCAN_TILE);
    uint32_t *tiles = (uint32_t *)ws;
    scan_reduce_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, tiles); D3D_LAUNCHED();
    scan_tiles_kernel<<<1, SCAN_THREADS, 0, st>>>(tiles, ntiles, total_out); D3D_LAUNCHED();
    scan_apply_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, tiles, out); D3D_LAUNCHED();
    return D3D_OK;
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits, (u64 key, u32 value) pairs.
// per pass: (1) per-tile digit histogram, (2) exclusive scan over the digit-major [256][ntiles] table,
// (3) stable scatter: rank inside the tile by warp match-any + per-warp digit counters.
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;                       // consecutive rounds of 32 keys per warp
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 2048 keys per block
constexpr int RS_BINS = 256;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, int64_t ntiles,
                                                             uint32_t *__restrict__ hist /*[256][ntiles]*/)
{
    __shared__ uint32_t h[RS_BINS];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        int64_t i = base + k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, int64_t n,
                                                                int shift, int64_t ntiles, const uint32_t *__restrict__ offs /*[256][ntiles] scanned*/,
                                                                uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
    __shared__ uint32_t cnt[RS_WARPS][RS_BINS];  // running per-warp digit counters -> per-warp totals
    __shared__ uint32_t binbase[RS_BINS];
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();

    const unsigned w = threadIdx.x >> 5, lane = lane_id();
    // warp w owns the contiguous chunk [w*32*ITEMS, (w+1)*32*ITEMS) of the tile: index order == (round, lane)
    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * 32 * RS_ITEMS;
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        int64_t i = wbase + k * 32 + lane;
        bool ok = i < n;
        key[k] = ok ? keys_in[i] : ~0ull;
        val[k] = ok ? vals_in[i] : 0u;
        unsigned d = (unsigned)(key[k] >> shift) & 0xff;
        unsigned act = __ballot_sync(0xffffffffu, ok);
        unsigned peers = __match_any_sync(0xffffffffu, ok ? d : 0x100u + lane) & act;  // inactive lanes match nobody
        uint32_t before = 0;
        if (ok) before = cnt[w][d];
        __syncwarp();
        if (ok) {
            rank[k] = before + __popc(peers & lanemask_lt());
            if ((peers & lanemask_lt()) == 0) cnt[w][d] = before + __popc(peers);  // leader of the peer group
        }
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: exclusive scan over warps of cnt[.][d], plus the global offset of (digit d, this tile)
        unsigned d = threadIdx.x;
        uint32_t run = offs[(int64_t)d * ntiles + blockIdx.x];
        binbase[d] = run;
        uint32_t acc = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { uint32_t c = cnt[ww][d]; cnt[ww][d] = acc; acc += c; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        int64_t i = wbase + k * 32 + lane;
        if (i < n) {
            unsigned d = (unsigned)(key[k] >> shift) & 0xff;
            uint32_t pos = binbase[d] + cnt[w][d] + rank[k];
            keys_out[pos] = key[k];
            vals_out[pos] = val[k];
        }
    }
}


// Small inputs (up to RS_COOP_TILES tiles = 131 072 keys: the NMS score sort, 50 000 boxes on C3): all passes in ONE cooperative launch.
// The five launches per pass of the general path (histogram, three scan kernels, scatter) are launch latency and nothing else at this
// size (C3: 40 launches, 0.2 ms for 600 KB of keys).  Here CTA t owns tile t for the whole sort; per pass it publishes its digit
// histogram, the grid synchronises, thread d of every CTA adds up digit d's column of the [256][ntiles] table itself (at most 64
// loads from L2: total of the digit and the part that belongs to earlier tiles) and a block scan over the 256 totals gives the
// digit bases -- no scan kernels -- then the tile scatters exactly as rs_scatter_kernel does, and the grid synchronises again
// before the next pass reads what this one wrote.  Same stable order, same digit ranks.
constexpr int RS_COOP_TILES = 64;

__global__ void __launch_bounds__(RS_THREADS) rs_coop_kernel(uint64_t *k0, uint32_t *v0, uint64_t *k1, uint32_t *v1, int64_t n, int passes, int ntiles,
                                                             uint32_t *__restrict__ hist /*[256][ntiles]*/, int shift0, const uint32_t *__restrict__ run_if)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    if (run_if && *run_if == 0u) {   // (grid-uniform) nothing to sort: the sequence the previous step left in k1 / v1 is the result
        for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_THREADS) { k0[i] = k1[i]; v0[i] = v1[i]; }
        return;
    }
    __shared__ uint32_t cnt[RS_WARPS][RS_BINS];
    __shared__ uint32_t binbase[RS_BINS];
    const unsigned w = threadIdx.x >> 5, lane = lane_id();
    const int tile = blockIdx.x;
    const int64_t wbase = (int64_t)tile * RS_TILE + (int64_t)w * 32 * RS_ITEMS;
    uint64_t *ki = k0, *ko = k1;
    uint32_t *vi = v0, *vo = v1;
    for (int p = 0; p < passes; p++) {
        const int shift = shift0 + p * 8;
        for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&cnt[0][0])[i] = 0;
        __syncthreads();
        // ranks inside the tile (warp match + per-warp digit counters), as in rs_scatter_kernel
        uint64_t key[RS_ITEMS];
        uint32_t val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
        for (int k = 0; k < RS_ITEMS; k++) {
            const int64_t i = wbase + k * 32 + lane;
            const bool ok = i < n;
            key[k] = ok ? ki[i] : ~0ull;
            val[k] = ok ? vi[i] : 0u;
            const unsigned d = (unsigned)(key[k] >> shift) & 0xff;
            const unsigned act = __ballot_sync(0xffffffffu, ok);
            const unsigned peers = __match_any_sync(0xffffffffu, ok ? d : 0x100u + lane) & act;
            uint32_t before = 0;
            if (ok) before = cnt[w][d];
            __syncwarp();
            if (ok) {
                rank[k] = before + __popc(peers & lanemask_lt());
                if ((peers & lanemask_lt()) == 0) cnt[w][d] = before + __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        {   // thread d: the tile's count of digit d (published), exclusive scan over the warps
            const unsigned d = threadIdx.x;
            uint32_t acc = 0;
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ww++) { const uint32_t c = cnt[ww][d]; cnt[ww][d] = acc; acc += c; }
            hist[(int64_t)d * ntiles + tile] = acc;
        }
        grid.sync();
        {   // thread d: digit d over all tiles, and over the tiles before this one
            const unsigned d = threadIdx.x;
            const uint32_t *col = hist + (int64_t)d * ntiles;
            uint32_t total = 0, before = 0;
            for (int t = 0; t < ntiles; t++) { const uint32_t c = __ldcg(col + t); if (t < tile) before += c; total += c; }
            uint32_t tot;
            binbase[d] = block_excl_scan(total, &tot) + before;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RS_ITEMS; k++) {
            const int64_t i = wbase + k * 32 + lane;
            if (i < n) {
                const unsigned d = (unsigned)(key[k] >> shift) & 0xff;
                const uint32_t pos = binbase[d] + cnt[w][d] + rank[k];
                ko[pos] = key[k];
                vo[pos] = val[k];
            }
        }
        grid.sync();   // the next pass reads this pass's output and rewrites the histogram table
        uint64_t *tk = ki; ki = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
    }
    if (ki != k0) {   // odd number of passes: the sorted sequence goes back to the caller's arrays
        for (int k = 0; k < RS_ITEMS; k++) {
            const int64_t i = wbase + k * 32 + lane;
            if (i < n) { k0[i] = ki[i]; v0[i] = vi[i]; }
        }
    }
}

size_t radix_sort_workspace_bytes(int64_t n)
{
    int64_t ntiles = cdiv(n > 0 ? n : 1, RS_TILE);
    size_t hist = align_up((size_t)RS_BINS * ntiles * sizeof(uint32_t));
    return align_up((size_t)n * sizeof(uint64_t)) + align_up((size_t)n * sizeof(uint32_t)) + hist + scan_workspace_bytes(RS_BINS * ntiles);
}

// Sorts in place semantically: on return keys/vals hold the sorted sequence (ping-pongs through the
// workspace; copies back when the pass count is odd).  key_bits: number of significant low bits.
int radix_sort_pairs_u64(uint64_t *keys, uint32_t *vals, int64_t n, int key_bits, void *ws, size_t ws_bytes, cudaStream_t st)
{
    if (n <= 1) return D3D_OK;
    if (n >= (1ll << 32)) return D3D_ERR_INVALID_ARGUMENT;
    if (ws_bytes < radix_sort_workspace_bytes(n)) return D3D_ERR_WORKSPACE;
    Arena a(ws, ws_bytes);
    uint64_t *k2 = a.take<uint64_t>(n);
    uint32_t *v2 = a.take<uint32_t>(n);
    int64_t ntiles = cdiv(n, RS_TILE);
    uint32_t *hist = a.take<uint32_t>((size_t)RS_BINS * ntiles);
    void *scan_ws = a.take<char>(scan_workspace_bytes(RS_BINS * ntiles));
    int passes = (key_bits + 7) / 8;
    if (passes < 1) passes = 1;
    if (ntiles <= RS_COOP_TILES && tuning(D3D_TUNE_SORT_COOP, 1)) {   // one CTA per tile: always co-resident (<= 64 CTAs of 256 threads)
        int nt = (int)ntiles, shift0 = 0;
        const uint32_t *run_if = nullptr;
        void *args[] = {&keys, &vals, &k2, &v2, &n, &passes, &nt, &hist, &shift0, &run_if};
        D3D_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)rs_coop_kernel, dim3((unsigned)ntiles), dim3(RS_THREADS), args, 0, st)); D3D_LAUNCHED();
        return D3D_OK;
    }
    uint64_t *ki = keys, *ko = k2;
    uint32_t *vi = vals, *vo = v2;
    for (int p = 0; p < passes; p++) {
        rs_hist_kernel<<<(unsigned)ntiles, RS_THREADS, 0, st>>>(ki, n, p * 8, ntiles, hist); D3D_LAUNCHED();
        int rc = exclusive_scan_u32(hist, hist, RS_BINS * ntiles, nullptr, scan_ws, st);
        if (rc) return rc;
        rs_scatter_kernel<<<(unsigned)ntiles, RS_THREADS, 0, st>>>(ki, vi, n, p * 8, ntiles, hist, ko, vo); D3D_LAUNCHED();
        uint64_t *tk = ki; ki = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
    }
    if (ki != keys) {
        D3D_CUDA_TRY(cudaMemcpyAsync(keys, ki, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        D3D_CUDA_TRY(cudaMemcpyAsync(vals, vi, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    return D3D_OK;
}

// ---- 64-bit keys whose low half rarely decides the order (scores in double precision: the high half holds sign, exponent and 20
// mantissa bits).  Four passes over the high half, then every element looks at the run of equal high halves around it -- one or two
// elements almost always -- and takes its place inside the run by counting (low half, then position: the stable order of the full key).
// A run longer than RS_RUN_MAX on either side raises a device flag and the eight passes of the full sort run on the sequence as it
// stands (a stable sort of any arrangement that kept equal keys in index order gives the same result); otherwise that launch only
// copies the fixed-up sequence back to the caller's arrays.  Same result as radix_sort_pairs_u64(..., 64, ...).
constexpr int RS_RUN_MAX = 16;

__global__ void __launch_bounds__(256) rs_runfix_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n,
                                                        uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t *__restrict__ flag)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint64_t k = keys[p];
    const uint32_t hi = (uint32_t)(k >> 32), lo = (uint32_t)k;
    int64_t s = p, e = p + 1;
    while (s > 0 && p - s < RS_RUN_MAX && (uint32_t)(keys[s - 1] >> 32) == hi) s--;
    while (e < n && e - p <= RS_RUN_MAX && (uint32_t)(keys[e] >> 32) == hi) e++;
    if ((s > 0 && (uint32_t)(keys[s - 1] >> 32) == hi) || (e < n && (uint32_t)(keys[e] >> 32) == hi)) *flag = 1u;   // the run goes on: the full sort decides
    int64_t at = s;
    for (int64_t q = s; q < e; q++) {
        const uint32_t lq = (uint32_t)keys[q];
        at += (lq < lo || (lq == lo && q < p)) ? 1 : 0;
    }
    keys_out[at] = k;
    vals_out[at] = vals[p];
}

int radix_sort_pairs_u64_hi32(uint64_t *keys, uint32_t *vals, int64_t n, void *ws, size_t ws_bytes, uint32_t *flag, cudaStream_t st)
{
    if (n <= 1) return D3D_OK;
    const int64_t ntiles = cdiv(n, RS_TILE);
    if (ntiles > RS_COOP_TILES || !tuning(D3D_TUNE_SORT_COOP, 1) || tuning(D3D_TUNE_SORT_COOP, 1) == 2)   // D3D_B200_SORT_COOP=2: all eight passes (A/B, tests)
        return radix_sort_pairs_u64(keys, vals, n, 64, ws, ws_bytes, st);
    if (ws_bytes < radix_sort_workspace_bytes(n)) return D3D_ERR_WORKSPACE;
    Arena a(ws, ws_bytes);
    uint64_t *k2 = a.take<uint64_t>(n);
    uint32_t *v2 = a.take<uint32_t>(n);
    uint32_t *hist = a.take<uint32_t>((size_t)RS_BINS * ntiles);
    D3D_CUDA_TRY(cudaMemsetAsync(flag, 0, 4, st));
    int64_t n_ = n;
    int nt = (int)ntiles;
    {
        int passes = 4, shift0 = 32;
        const uint32_t *run_if = nullptr;
        void *args[] = {&keys, &vals, &k2, &v2, &n_, &passes, &nt, &hist, &shift0, &run_if};
        D3D_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)rs_coop_kernel, dim3((unsigned)ntiles), dim3(RS_THREADS), args, 0, st)); D3D_LAUNCHED();
    }
    rs_runfix_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(keys, vals, n, k2, v2, flag); D3D_LAUNCHED();
    {
        int passes = 8, shift0 = 0;
        const uint32_t *run_if = flag;
        void *args[] = {&keys, &vals, &k2, &v2, &n_, &passes, &nt, &hist, &shift0, &run_if};
        D3D_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)rs_coop_kernel, dim3((unsigned)ntiles), dim3(RS_THREADS), args, 0, st)); D3D_LAUNCHED();
    }
    return D3D_OK;
}

}  // namespace d3d
