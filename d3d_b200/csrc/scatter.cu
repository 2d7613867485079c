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
#include <cuda.h>
#include <string.h>
#include "prims.cuh"

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
// one axis: lower / upper neighbour and their weights.  Mirrors _fill_lcoords (scatter.cpp:34-77) including its quirk that an
// integral in-range coordinate gives both coincident neighbours weight 1.
template <typename T>
__device__ __forceinline__ void axis_neighbours(T x, int dmax, int *lo, int *hi, T *wlo, T *whi)
{
    if (x > (T)dmax) { *lo = *hi = dmax; *wlo = *whi = T(0.5); }
    else if (x < (T)0) { *lo = *hi = 0; *wlo = *whi = T(0.5); }
    else {
        int k = (int)x;            // truncation; x >= 0 here so this is floor except for the checks below
        int fl = k > x ? k - 1 : k;   // _floor, scatter.cpp:22-27
        int ce = k < x ? k + 1 : k;   // _ceil,  scatter.cpp:28-33
        *lo = fl; *hi = ce;
        *wlo = add_rn<T>(add_rn<T>(T(1), -x), (T)fl);   // 1 - x + floor
        *whi = add_rn<T>(add_rn<T>(T(1), x), -(T)ce);   // 1 + x - ceil
    }
}

// neighbour j (bit d of j = take the upper neighbour along dim d): linear offset in the plane and weight.
template <typename T, int DIM, bool LINEAR>
__device__ __forceinline__ void neighbours(const T *__restrict__ crd, const ScDims &dm, long long off[1 << DIM], T wgt[1 << DIM])
{
    int lo[DIM], hi[DIM]; T wlo[DIM], whi[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) axis_neighbours<T>(crd[d + 1], dm.d[d] - 1, &lo[d], &hi[d], &wlo[d], &whi[d]);
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

// ------------------------------------------------------------------ tile path (2-D maps, fp32)
// The gather above costs one 32-byte sector per 4-byte neighbour and channel plane (lanes of a warp are 32 channels: 32 planes, 2.25 MB
// apart on the C2s map), 3x the algorithmic bytes.  When the points are dense enough that this exceeds the map itself, the map is
// streamed instead: the points are binned by the 6 x 64-cell tile of their lower neighbour, and one CTA per (tile, 32-channel chunk,
// slice of <= 256 points) stages the tile's 7 x 65 cells (one halo row and column for the upper neighbours) of its 32 channel planes
// in shared memory with 4-byte cp.async -- row segments of 65 consecutive floats, full sectors -- and serves its points from there:
// a warp per point, a lane per channel, plane pitch 455 words (odd: the 32 lanes hit 32 banks; 476 with the 68-cell rows of a TMA box: four lanes per bank), output rows written as 128-byte
// lines.  Tiles with few points skip the staging and gather from global memory.  The backward pass accumulates the tile in shared
// memory (atomicAdd on shared memory) and sends each non-zero cell to the map gradient once, as a RED on consecutive addresses.
// Arithmetic per (point, channel) is the gather kernel's, operation by operation: forward outputs are bit-identical on both paths.
#ifndef D3D_ST_H
#define D3D_ST_H 6
#endif
constexpr int ST_H = D3D_ST_H, ST_W = 64, ST_CH = 32, ST_ROWS = ST_H + 1, ST_COLS = ST_W + 1;
constexpr int ST_TMA_COLS = 68;   // TMA boxes are multiples of 16 bytes wide: 65 cells + 3
template <bool TMA> struct StLayout { static constexpr int PITCH = TMA ? ST_TMA_COLS : ST_COLS, PLANE = ST_ROWS * PITCH; };   // 476 / 455 words per channel plane
constexpr int ST_SLICE = 256;      // points per work item
constexpr int ST_DIRECT = 12;      // tiles with fewer points gather from global memory (12 points x 2 rows x 32 B < the 7 x 288 B of a staged plane)
constexpr int ST_THREADS = 256;
constexpr int64_t ST_MAX_TILES = 1 << 16;

struct StGeom { int H, W, nty, ntx; int64_t ntiles, nbatch; };

constexpr int ST_HIST = 8192;   // tiles whose counters fit the CTA-level histogram of the count kernel
constexpr int ST_CPT = 4;       // points per thread of the count kernel
template <typename T>
__device__ __forceinline__ uint32_t st_tile_of(const T *__restrict__ coord, int64_t i, const StGeom &g)
{
    const T *crd = coord + i * 3;
    const int b = (int)crd[0];
    if (b < 0 || b >= g.nbatch) return 0xffffffffu;   // the gather path would read outside the map: the point is skipped
    int lo0, hi0, lo1, hi1; T w0, w1, w2, w3;
    axis_neighbours<T>(crd[1], g.H - 1, &lo0, &hi0, &w0, &w1);
    axis_neighbours<T>(crd[2], g.W - 1, &lo1, &hi1, &w2, &w3);
    return (uint32_t)(((int64_t)b * g.nty + lo0 / ST_H) * g.ntx + lo1 / ST_W);
}

// tile and arrival rank of every point.  The tile under the sensor holds a sixth of a lidar frame and same-address global atomics
// serialise, so a CTA first ranks its 1024 points per tile in a shared-memory histogram and reserves one range per (CTA, tile).
template <typename T, bool HIST>
__global__ void __launch_bounds__(256) st_count_kernel(const T *__restrict__ coord, int64_t n, StGeom g, uint32_t *__restrict__ cnt, uint2 *__restrict__ tile_rank)
{
    extern __shared__ uint32_t st_hist[];
    const int64_t i0 = (int64_t)blockIdx.x * (256 * ST_CPT) + threadIdx.x;
    uint32_t t[ST_CPT], r[ST_CPT];
    if (HIST) {
        for (int e = threadIdx.x; e < (int)g.ntiles; e += 256) st_hist[e] = 0;
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < ST_CPT; u++) {
        const int64_t i = i0 + u * 256;
        t[u] = i < n ? st_tile_of<T>(coord, i, g) : 0xffffffffu;
        r[u] = 0;
        if (HIST) { if (t[u] != 0xffffffffu) r[u] = atomicAdd(st_hist + t[u], 1u); }
    }
    if (HIST) {
        __syncthreads();
        for (int e = threadIdx.x; e < (int)g.ntiles; e += 256) { const uint32_t c = st_hist[e]; if (c) st_hist[e] = atomicAdd(cnt + e, c); }   // count -> base of this CTA's range
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < ST_CPT; u++) {
        const int64_t i = i0 + u * 256;
        if (!HIST) {   // one global atomic per distinct tile of the warp
            const unsigned peers = __match_any_sync(0xffffffffu, t[u]), lane = threadIdx.x & 31;
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if ((int)lane == leader && t[u] != 0xffffffffu) base = atomicAdd(cnt + t[u], (uint32_t)__popc(peers));
            r[u] = __shfl_sync(0xffffffffu, base, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        } else if (t[u] != 0xffffffffu) r[u] += st_hist[t[u]];
        if (i < n) tile_rank[i] = make_uint2(t[u], r[u]);
    }
}

// one CTA: exclusive prefix of the tile counts (ptr) and the work-item list: tile t yields ceil(cnt / ST_SLICE) items (tile, point range, tile count)
__global__ void __launch_bounds__(1024) st_plan_kernel(const uint32_t *__restrict__ cnt, int64_t ntiles, uint32_t *__restrict__ ptr, uint4 *__restrict__ items,
                                                       uint32_t *__restrict__ nitems)
{
    __shared__ uint32_t wsum[32][2];
    __shared__ uint32_t carry[2];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
    __syncthreads();
    for (int64_t base = 0; base < ntiles; base += 1024) {
        const int64_t t = base + threadIdx.x;
        const uint32_t c = t < ntiles ? cnt[t] : 0u, it = (c + ST_SLICE - 1) / ST_SLICE;
        uint32_t pc = c, pi = it;   // inclusive warp prefixes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, pc, d), b = __shfl_up_sync(0xffffffffu, pi, d);
            if ((int)lane >= d) { pc += a; pi += b; }
        }
        if (lane == 31) { wsum[warp][0] = pc; wsum[warp][1] = pi; }
        __syncthreads();
        uint32_t oc = carry[0], oi = carry[1];
        for (unsigned w = 0; w < warp; w++) { oc += wsum[w][0]; oi += wsum[w][1]; }
        const uint32_t ec = oc + pc - c, ei = oi + pi - it;
        if (t < ntiles) {
            ptr[t] = ec;
            for (uint32_t k = 0; k < it; k++) items[ei + k] = make_uint4((uint32_t)t, ec + k * ST_SLICE, min(ec + c, ec + (k + 1) * ST_SLICE), c);   // tile, first point, end, points of the tile
        }
        __syncthreads();
        if (threadIdx.x == 1023) { carry[0] = oc + pc; carry[1] = oi + pi; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { ptr[ntiles] = carry[0]; *nitems = carry[1]; }
}

template <typename T>
__global__ void __launch_bounds__(256) st_place_kernel(const T *__restrict__ coord, int64_t n, const uint2 *__restrict__ tile_rank, const uint32_t *__restrict__ ptr,
                                                       float4 *__restrict__ sorted)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 tr = tile_rank[i];
    // the record carries the coordinates: the tile kernel reads one 16-byte record per point instead of chasing index -> coordinates
    if (tr.x != 0xffffffffu) sorted[ptr[tr.x] + tr.y] = make_float4(__uint_as_float((uint32_t)i), (float)coord[i * 3 + 1], (float)coord[i * 3 + 2], 0.f);
}

// TMA / mbarrier primitives of the tile kernel (sm_100a PTX)
__device__ __forceinline__ void st_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void st_mbar_expect(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ST_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ST_DONE;\n"
        "bra ST_WAIT;\n"
        "ST_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// box {ST_TMA_COLS, ST_ROWS, ST_CH} of the map [planes][H][W] -> shared memory, cells outside the map arrive as zeros
__device__ __forceinline__ void st_tma_load(uint32_t dst, const CUtensorMap *tm, int x, int y, int z, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
// map box += shared-memory box (element-wise add performed by the memory system; cells outside the map are dropped)
__device__ __forceinline__ void st_tma_reduce_add(const CUtensorMap *tm, int x, int y, int z, uint32_t src)
{
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(x), "r"(y), "r"(z) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory is read before the CTA leaves
}

// FWD: out[i, c] = sum_j w_j * image[b, c, neighbour_j(i)];  !FWD: image_grad[b, c, neighbour_j(i)] += w_j * grad[i, c]
// TMA: the tile moves with one tensor copy (forward) / one tensor reduce-add (backward) per CTA; otherwise (map rows that are no
// multiple of 16 bytes) with 4-byte cp.async per cell and REDs per non-zero cell.
template <typename T, bool LINEAR, bool FWD, bool TMA>
__global__ void __launch_bounds__(ST_THREADS) st_tile_kernel(const T *__restrict__ src, int64_t nchan, StGeom g, const float4 *__restrict__ sorted,
                                                             const uint4 *__restrict__ items, const uint32_t *__restrict__ nitems, T *__restrict__ dst,
                                                             const __grid_constant__ CUtensorMap tmap)
{
    constexpr int PITCH = StLayout<TMA>::PITCH, PLANE = StLayout<TMA>::PLANE;
    extern __shared__ float st_smem_raw[];
    __shared__ __align__(8) uint64_t st_bar;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x >= *nitems) return;
    // the tile starts at the first 128-byte boundary of the dynamic shared memory (TMA destination alignment)
    const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(st_smem_raw), tile_s = (raw_s + 127u) & ~127u;
    T *tile = reinterpret_cast<T *>(reinterpret_cast<char *>(st_smem_raw) + (tile_s - raw_s));   // [ST_CH][ST_ROWS][PITCH]
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&st_bar);
    const uint4 it = items[blockIdx.x];
    const int64_t t = it.x;
    const int tx = (int)(t % g.ntx), ty = (int)((t / g.ntx) % g.nty);
    const int64_t b = t / ((int64_t)g.ntx * g.nty);
    const uint32_t p0 = it.y, p1 = it.z, tcnt = it.w;
    const int c0 = blockIdx.y * ST_CH, nc = (int)min((int64_t)ST_CH, nchan - c0);
    const int y0 = ty * ST_H, x0 = tx * ST_W;
    const int rows = min(ST_ROWS, g.H - y0), cols = min(ST_COLS, g.W - x0);
    const int64_t plane = (int64_t)g.H * g.W;
    const int zplane = (int)(b * nchan + c0);
    const bool direct = tcnt < (uint32_t)ST_DIRECT;   // the same for every slice and thread of the tile
    if (!direct) {
        if (FWD && TMA) {
            if (threadIdx.x == 0) {
                st_mbar_init(bar_s, 1);
                st_mbar_expect(bar_s, (uint32_t)(ST_CH * PLANE * sizeof(T)));
                st_tma_load(tile_s, &tmap, x0, y0, zplane, bar_s);
            }
        } else if (FWD) {
            const T *base = src + ((int64_t)b * nchan + c0) * plane + (int64_t)y0 * g.W + x0;
            // a warp per (channel, row), consecutive lanes on consecutive columns: full sectors from global memory, consecutive banks in shared
            // memory.  The warps step through the rows 8 at a time with running (channel, row) counters: no division inside the loop.
            const int step_c = (ST_THREADS / 32) / rows, step_r = (ST_THREADS / 32) % rows;
            for (int cc = (int)warp / rows, r = (int)warp % rows; cc < nc;) {
                const T *grow = base + (int64_t)cc * plane + (int64_t)r * g.W + lane;
                const uint32_t srow = tile_s + (uint32_t)((cc * PLANE + r * PITCH + (int)lane) * (int)sizeof(T));
                if ((int)lane < cols) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(srow), "l"(grow) : "memory");
                if ((int)lane + 32 < cols) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(srow + 128u), "l"(grow + 32) : "memory");
                if ((int)lane + 64 < cols) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(srow + 256u), "l"(grow + 64) : "memory");
                cc += step_c; r += step_r;
                if (r >= rows) { r -= rows; cc++; }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            float4 *t4 = reinterpret_cast<float4 *>(tile);   // 128-byte aligned, ST_CH * PLANE is a multiple of 4 for both layouts' element counts rounded below
            for (int e = threadIdx.x; e < ST_CH * PLANE / 4; e += ST_THREADS) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = ST_CH * PLANE / 4 * 4 + threadIdx.x; e < ST_CH * PLANE; e += ST_THREADS) tile[e] = T(0);
        }
    }
    // the points of the slice are dealt round robin to the 8 warps (a typical tile holds 32 points: 4 per warp instead of all 32 on one);
    // lane l prepares the l-th point of its warp's share, the warp then serves the prepared points one after the other
    const int c = c0 + (int)lane;
    const bool chan_ok = (int)lane < nc;
    uint32_t myi = 0;
    int lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
    T wy0 = T(0), wy1 = T(0), wx0 = T(0), wx1 = T(0);
    constexpr uint32_t NW = ST_THREADS / 32;
    auto prepare = [&](uint32_t q) {
        if (q + lane * NW < p1) {
            const float4 rec = sorted[q + lane * NW];
            myi = __float_as_uint(rec.x);
            axis_neighbours<T>((T)rec.y, g.H - 1, &lo0, &hi0, &wy0, &wy1);
            axis_neighbours<T>((T)rec.z, g.W - 1, &lo1, &hi1, &wx0, &wx1);
        }
    };
    uint32_t q0 = p0 + warp;   // this warp's points: q0, q0 + NW, q0 + 2 NW, ...
    prepare(q0);   // the first round's coordinates travel while the staging lands
    if (!direct) {
        if (FWD && !TMA) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();   // every thread of the CTA: the tile is staged / zeroed, or (TMA forward) the barrier is initialised
        if (FWD && TMA) st_mbar_wait(bar_s, 0);
    }
    for (; q0 < p1; q0 += ST_THREADS, prepare(q0)) {
        const int np = (int)min(32u, (p1 - q0 + NW - 1) / NW);
        for (int k0 = 0; k0 < np; k0 += 4) {   // four points per round: their gradient rows (backward) are in flight together
            uint32_t pi[4];
            T gr[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                pi[u] = __shfl_sync(0xffffffffu, myi, min(k0 + u, np - 1));
                gr[u] = T(0);
                if (!FWD && chan_ok && k0 + u < np) gr[u] = src[(int64_t)pi[u] * nchan + c];
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int k = min(k0 + u, np - 1);   // np is the same in every lane: the shuffles are converged
                const int a0 = __shfl_sync(0xffffffffu, lo0, k), a1 = __shfl_sync(0xffffffffu, hi0, k);
                const int b0 = __shfl_sync(0xffffffffu, lo1, k), b1 = __shfl_sync(0xffffffffu, hi1, k);
                const T u0 = __shfl_sync(0xffffffffu, wy0, k), u1 = __shfl_sync(0xffffffffu, wy1, k);
                const T v0 = __shfl_sync(0xffffffffu, wx0, k), v1 = __shfl_sync(0xffffffffu, wx1, k);
                if (!chan_ok || k0 + u >= np) continue;
                // neighbour j: bit 0 of j = upper neighbour along the first axis, bit 1 = along the second (the order of neighbours())
                T w[4];
                if (LINEAR) { w[0] = mul_rn<T>(mul_rn<T>(T(1), u0), v0); w[1] = mul_rn<T>(mul_rn<T>(T(1), u1), v0); w[2] = mul_rn<T>(mul_rn<T>(T(1), u0), v1); w[3] = mul_rn<T>(mul_rn<T>(T(1), u1), v1); }
                if (FWD) {
                    T v[4];
                    if (direct) {
                        const T *pl = src + ((int64_t)b * nchan + c) * plane;
                        v[0] = __ldg(pl + (int64_t)a0 * g.W + b0); v[1] = __ldg(pl + (int64_t)a1 * g.W + b0);
                        v[2] = __ldg(pl + (int64_t)a0 * g.W + b1); v[3] = __ldg(pl + (int64_t)a1 * g.W + b1);
                    } else {
                        const T *pl = tile + (int)lane * PLANE;
                        v[0] = pl[(a0 - y0) * PITCH + (b0 - x0)]; v[1] = pl[(a1 - y0) * PITCH + (b0 - x0)];
                        v[2] = pl[(a0 - y0) * PITCH + (b1 - x0)]; v[3] = pl[(a1 - y0) * PITCH + (b1 - x0)];
                    }
                    T sum = T(0);
#pragma unroll
                    for (int j = 0; j < 4; j++) sum = add_rn<T>(sum, LINEAR ? mul_rn<T>(v[j], w[j]) : v[j]);
                    dst[(int64_t)pi[u] * nchan + c] = LINEAR ? sum : sum / T(4);
                } else {
                    T a[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) a[j] = LINEAR ? mul_rn<T>(gr[u], w[j]) : gr[u] / T(4);
                    if (direct) {
                        T *pl = dst + ((int64_t)b * nchan + c) * plane;
                        atomicAdd(pl + (int64_t)a0 * g.W + b0, a[0]); atomicAdd(pl + (int64_t)a1 * g.W + b0, a[1]);
                        atomicAdd(pl + (int64_t)a0 * g.W + b1, a[2]); atomicAdd(pl + (int64_t)a1 * g.W + b1, a[3]);
                    } else {
                        T *pl = tile + (int)lane * PLANE;
                        atomicAdd(pl + (a0 - y0) * PITCH + (b0 - x0), a[0]); atomicAdd(pl + (a1 - y0) * PITCH + (b0 - x0), a[1]);
                        atomicAdd(pl + (a0 - y0) * PITCH + (b1 - x0), a[2]); atomicAdd(pl + (a1 - y0) * PITCH + (b1 - x0), a[3]);
                    }
                }
            }
        }
    }
    if (FWD || direct) return;
    if (TMA) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the accumulation becomes visible to the copy engine
        __syncthreads();
        // channels past nc and cells past the 65th column hold zeros; rows and columns past the map are dropped by the engine
        if (threadIdx.x == 0) st_tma_reduce_add(&tmap, x0, y0, zplane, tile_s);
        return;
    }
    __syncthreads();   // every thread of the CTA: the accumulation is complete
    T *base = dst + ((int64_t)b * nchan + c0) * plane + (int64_t)y0 * g.W + x0;
    const int step_c = (ST_THREADS / 32) / rows, step_r = (ST_THREADS / 32) % rows;
    for (int cc = (int)warp / rows, r = (int)warp % rows; cc < nc;) {   // a warp per (channel, row): REDs on consecutive addresses
        T *grow = base + (int64_t)cc * plane + (int64_t)r * g.W + lane;
        const T *srow = tile + cc * PLANE + r * PITCH + lane;
#pragma unroll
        for (int u = 0; u < 3; u++) {
            if ((int)lane + 32 * u < cols) {
                const T v = srow[32 * u];
                if (v != T(0)) atomicAdd(grow + 32 * u, v);
            }
        }
        cc += step_c; r += step_r;
        if (r >= rows) { r -= rows; cc++; }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*StEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static StEncodeFn st_encode_fn()
{
    static StEncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
        return (StEncodeFn)p;
    }();
    return fn;
}
// tensor map of a [planes][H][W] fp32 map with the tile box; false where TMA does not apply (rows no multiple of 16 bytes, unaligned base)
static bool st_tensor_map(const void *base, int64_t planes, const StGeom &g, CUtensorMap *tm)
{
    StEncodeFn enc = st_encode_fn();
    if (!enc || g.W % 4 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || planes <= 0 || planes > 0x7fffffffll) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)g.W * 4, (cuuint64_t)g.W * g.H * 4};
    const cuuint32_t box[3] = {ST_TMA_COLS, ST_ROWS, ST_CH}, estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool st_geom(int64_t nbatch, const ScDims &dm, StGeom *g)
{
    g->H = dm.d[0]; g->W = dm.d[1]; g->nbatch = nbatch;
    g->nty = (int)cdiv((int64_t)g->H, (int64_t)ST_H); g->ntx = (int)cdiv((int64_t)g->W, (int64_t)ST_W);
    g->ntiles = nbatch * g->nty * g->ntx;
    return g->ntiles > 0 && g->ntiles <= ST_MAX_TILES;
}
static size_t st_ws_bytes(int64_t n, int64_t ntiles)
{
    const size_t n1 = (size_t)(n > 0 ? n : 1), items = (size_t)ntiles + n1 / ST_SLICE + 1;
    return align_up(((size_t)ntiles + 1) * 4) * 2 + align_up(64) + align_up(n1 * sizeof(uint2)) + align_up(n1 * sizeof(float4)) + align_up(items * sizeof(uint4)) + 256;
}

template <typename T>
static int scatter_tiles(bool fwd, const T *coord, int64_t n, const T *src, int64_t nbatch, int64_t nchan, const ScDims &dm, int align, T *dst, void *ws, size_t ws_bytes,
                         bool reuse_plan, cudaStream_t st)
{
    StGeom g;
    if (!st_geom(nbatch, dm, &g) || n >= (1ll << 31)) return D3D_ERR_INVALID_ARGUMENT;
    if (!ws || ws_bytes < st_ws_bytes(n, g.ntiles)) return D3D_ERR_WORKSPACE;
    Arena a(ws, ws_bytes);
    uint32_t *cnt = a.take<uint32_t>((size_t)g.ntiles + 1), *ptr = a.take<uint32_t>((size_t)g.ntiles + 1), *nitems = a.take<uint32_t>(16);
    uint2 *tile_rank = a.take<uint2>(n);
    float4 *sorted = a.take<float4>(n);
    const int64_t max_items = g.ntiles + n / ST_SLICE + 1;
    uint4 *items = a.take<uint4>(max_items);
    if (!a.ok()) return D3D_ERR_WORKSPACE;
    const int64_t chunks = cdiv(nchan, (int64_t)ST_CH);
    if (max_items > 0x7fffffffll || chunks > 65535) return D3D_ERR_INVALID_ARGUMENT;
    if (!reuse_plan) {
        D3D_CUDA_TRY(cudaMemsetAsync(cnt, 0, ((size_t)g.ntiles + 1) * 4, st));
        const unsigned gc = (unsigned)cdiv(n, (int64_t)256 * ST_CPT);
        if (g.ntiles <= ST_HIST) st_count_kernel<T, true><<<gc, 256, (size_t)g.ntiles * 4, st>>>(coord, n, g, cnt, tile_rank);
        else st_count_kernel<T, false><<<gc, 256, 0, st>>>(coord, n, g, cnt, tile_rank);
        D3D_LAUNCHED();
        st_plan_kernel<<<1, 1024, 0, st>>>(cnt, g.ntiles, ptr, items, nitems); D3D_LAUNCHED();
        st_place_kernel<T><<<(unsigned)cdiv(n, 256), 256, 0, st>>>(coord, n, tile_rank, ptr, sorted); D3D_LAUNCHED();
    }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const bool tma = sizeof(T) == 4 && st_tensor_map(fwd ? (const void *)src : (const void *)dst, nbatch * nchan, g, &tmap);
    const dim3 grid((unsigned)max_items, (unsigned)chunks);
#define D3D_ST_LAUNCH(LIN, FWD, TMA)                                                                                                         \
    do {                                                                                                                                     \
        const size_t smem = (size_t)ST_CH * StLayout<TMA>::PLANE * sizeof(T) + 128;                                                          \
        static bool attr = false; /* the same value every time: a benign race */                                                             \
        if (!attr) { D3D_CUDA_TRY(cudaFuncSetAttribute(st_tile_kernel<T, LIN, FWD, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; } \
        st_tile_kernel<T, LIN, FWD, TMA><<<grid, ST_THREADS, smem, st>>>(src, nchan, g, sorted, items, nitems, dst, tmap);      \
    } while (0)
#define D3D_ST_LAUNCH2(LIN, FWD) do { if (tma) D3D_ST_LAUNCH(LIN, FWD, true); else D3D_ST_LAUNCH(LIN, FWD, false); } while (0)
    if (fwd) {
        // points with a batch index outside the map are skipped (the gather path, like the reference, would read outside the map): their rows stay unwritten
        if (align == D3D_ALIGN_LINEAR) D3D_ST_LAUNCH2(true, true); else D3D_ST_LAUNCH2(false, true);
    } else {
        if (align == D3D_ALIGN_LINEAR) D3D_ST_LAUNCH2(true, false); else D3D_ST_LAUNCH2(false, false);
    }
#undef D3D_ST_LAUNCH2
#undef D3D_ST_LAUNCH
    D3D_LAUNCHED();
    return D3D_OK;
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

static bool scatter_dims(int32_t dim, const int64_t *dims, ScDims *dm)
{
    if (dim < 1 || dim > 3 || !dims) return false;
    dm->plane = 1;
    for (int d = 0; d < 3; d++) {
        dm->d[d] = d < dim ? (int)dims[d] : 1;
        if (d < dim) { if (dims[d] <= 0 || dims[d] > 0x7fffffff) return false; dm->plane *= dims[d]; }
    }
    return true;
}

// tile path: 2-D fp32 maps whose points are dense enough that one 32-byte sector per neighbour row and channel (the gather's cost)
// exceeds the map itself; D3D_B200_SCATTER_PATH=gather|tiles forces one (tiles only where it applies)
static bool scatter_use_tiles(int64_t n, int32_t dim, int dtype, int64_t nbatch, const ScDims &dm, const void *ws)
{
    StGeom g;
    if (dim != 2 || dtype != D3D_F32 || !ws || n <= 0 || !st_geom(nbatch, dm, &g)) return false;
    const int mode = tuning(D3D_TUNE_SCATTER_PATH, 0);   // 1 gather, 2 tiles
    if (mode) return mode == 2;
    return n * 16 >= (int64_t)dm.plane * nbatch;
}

template <typename T>
static int scatter_impl(bool fwd, const void *coord, int64_t n, int32_t dim, const void *src, int64_t nbatch, int64_t nchan, const int64_t *dims, int align, void *dst,
                        void *ws, size_t ws_bytes, bool reuse_plan, cudaStream_t st)
{
    if (align != D3D_ALIGN_MEAN && align != D3D_ALIGN_LINEAR) return D3D_ERR_INVALID_ARGUMENT;   // reference: "Unsupported align type!"
    if (dim < 1 || dim > 3) return D3D_ERR_INVALID_ARGUMENT;                                    // reference: "Unsupported dimension size"
    if (n < 0 || nbatch < 0 || nchan < 0 || !dims) return D3D_ERR_INVALID_ARGUMENT;
    if (n * nchan > 0 && (!coord || !src || !dst)) return D3D_ERR_INVALID_ARGUMENT;
    ScDims dm;
    if (!scatter_dims(dim, dims, &dm)) return D3D_ERR_INVALID_ARGUMENT;
    if (n * nchan >= (1ll << 31) * 256) return D3D_ERR_INVALID_ARGUMENT;
    if (n * nchan > 0 && sizeof(T) == 4 && scatter_use_tiles(n, dim, D3D_F32, nbatch, dm, ws))
        return scatter_tiles<float>(fwd, (const float *)coord, n, (const float *)src, nbatch, nchan, dm, align, (float *)dst, ws, ws_bytes, reuse_plan, st);
    switch (dim) {
    case 1: return scatter_launch<T, 1>(fwd, (const T *)coord, n, (const T *)src, nchan, dm, align, (T *)dst, st);
    case 2: return scatter_launch<T, 2>(fwd, (const T *)coord, n, (const T *)src, nchan, dm, align, (T *)dst, st);
    default: return scatter_launch<T, 3>(fwd, (const T *)coord, n, (const T *)src, nchan, dm, align, (T *)dst, st);
    }
}

}  // namespace d3d

using namespace d3d;
extern "C" size_t d3d_aligned_scatter_workspace_bytes(int64_t n, int32_t dim, int64_t nbatch, const int64_t *dims_host)
{
    ScDims dm;
    StGeom g;
    if (dim != 2 || n <= 0 || nbatch <= 0 || !scatter_dims(dim, dims_host, &dm) || !st_geom(nbatch, dm, &g)) return 0;   // the gather path needs none
    return st_ws_bytes(n, g.ntiles);
}
extern "C" int d3d_aligned_scatter_forward_ws(const void *coord, int64_t n, int32_t dim, const void *image, int64_t nbatch, int64_t nchan, const int64_t *dims_host,
                                              int align, int dtype, void *out, void *workspace, size_t workspace_bytes, int reuse_plan, void *stream)
{
    return dtype == D3D_F64 ? scatter_impl<double>(true, coord, n, dim, image, nbatch, nchan, dims_host, align, out, nullptr, 0, false, (cudaStream_t)stream)
                            : scatter_impl<float>(true, coord, n, dim, image, nbatch, nchan, dims_host, align, out, workspace, workspace_bytes, reuse_plan != 0, (cudaStream_t)stream);
}
extern "C" int d3d_aligned_scatter_backward_ws(const void *coord, int64_t n, int32_t dim, const void *grad, int64_t nbatch, int64_t nchan, const int64_t *dims_host,
                                               int align, int dtype, void *image_grad, void *workspace, size_t workspace_bytes, int reuse_plan, void *stream)
{
    return dtype == D3D_F64 ? scatter_impl<double>(false, coord, n, dim, grad, nbatch, nchan, dims_host, align, image_grad, nullptr, 0, false, (cudaStream_t)stream)
                            : scatter_impl<float>(false, coord, n, dim, grad, nbatch, nchan, dims_host, align, image_grad, workspace, workspace_bytes, reuse_plan != 0, (cudaStream_t)stream);
}
extern "C" int d3d_aligned_scatter_forward(const void *coord, int64_t n, int32_t dim, const void *image, int64_t nbatch, int64_t nchan, const int64_t *dims_host, int align,
                                           int dtype, void *out, void *stream)
{ return d3d_aligned_scatter_forward_ws(coord, n, dim, image, nbatch, nchan, dims_host, align, dtype, out, nullptr, 0, 0, stream); }
extern "C" int d3d_aligned_scatter_backward(const void *coord, int64_t n, int32_t dim, const void *grad, int64_t nbatch, int64_t nchan, const int64_t *dims_host, int align,
                                            int dtype, void *image_grad, void *stream)
{ return d3d_aligned_scatter_backward_ws(coord, n, dim, grad, nbatch, nchan, dims_host, align, dtype, image_grad, nullptr, 0, 0, stream); }
