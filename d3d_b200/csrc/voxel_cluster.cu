// voxel_cluster.cu -- voxelization fast path: one thread-block CLUSTER per frame, one persistent launch.
//
// Replaces the same reference functions as voxel.cu (voxelize_sparse + voxelize_filter and
// voxelize_3d_dense, d3d/voxel/voxelize.cpp:288-484 and :45-199) for the common configurations; voxel.cu's
// sort-based pipeline stays as the general fallback (DESCENDING filter, dense reductions, > 2^32 cells).
//
// Why a cluster per frame: frames are independent, a frame's working set (120k-200k points) is a few MB,
// and the reference's sequential semantics (voxel ids in order of first appearance, first max_points points
// of a voxel in input order) need frame-wide prefix sums.  A cluster of 8 CTAs x 1024 threads owns a frame
// from the first point load to the last output store: the phases are separated by cluster barriers
// (barrier.cluster, ~1 us) instead of kernel launches, prefix sums cross CTAs through distributed shared
// memory, and the whole scratch state of the frames in flight (hash table, per-point words) stays in L2, so
// HBM sees the algorithmic traffic only: 16 B/point in, 32 B/kept point + 28 B/voxel out.
//
// Determinism without a sort.  Everything that decides an output is an order-independent function of the
// point set, although the hash-table slots themselves are handed out in race order:
//   * a voxel's slot holds (cell key << 32 | smallest point index) maintained by atomicMin, and its point
//     count by atomicAdd -- commutative, so first[] and count[] are unique;
//   * voxel ids = exclusive prefix sum, in point order, of "this point is the first of a voxel that passes
//     the voxel filter" -- exactly the reference's first-appearance numbering;
//   * "first max_points points of a voxel" only needs order inside voxels that hold MORE than max_points
//     points (a few per cent): their point indices are appended to a per-voxel list (position = the value
//     atomicAdd returned, any order) and a point's rank is the number of smaller indices in that list, found
//     by a direct scan (<= 32 entries) or, for big voxels, by a warp that extracts the max_points smallest
//     indices one minimum at a time -- both independent of the list order;
//   * kept points are compacted with a second prefix sum in point order.
#include "voxel.cuh"
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace d3d {

constexpr int VC_THREADS = 1024;
constexpr int VC_WARPS = VC_THREADS / 32;
constexpr int VC_MAX_CSIZE = 16;
constexpr int VC_MAX_CLUSTERS = 74;            // frames in flight the workspace is sized for
constexpr uint32_t VC_NONE = 0xffffffffu;
constexpr unsigned long long VC_EMPTY = ~0ull;
// per-point word after P3: VC_NONE (dropped) | VC_SINGLE + provisional voxel id (the only point of its voxel: no
// further table access needed) | table slot
constexpr uint32_t VC_SINGLE = 1u << 31;
// list record of a crowded voxel: [0] provisional voxel id, [1] big voxels: keep indices below this, [2] fill cursor, [3..] point indices
constexpr int VC_LIST_HDR = 3;
constexpr int VC_LOCAL_BITS = 20;              // provisional voxel id = (warp << 20) | rank inside the warp's chunk

struct __align__(16) VcEntry {
    unsigned long long kf;   // (cell key << 32) | smallest frame-local index of a point in the cell
    uint32_t count;          // points in the cell beyond the one that claimed the slot (total - 1)
    uint32_t aux;            // crowded voxel: start of its list record; otherwise the provisional voxel id
};

struct VcLayout {            // byte offsets inside one cluster's workspace slice
    size_t tab, pw, pkey, cw, lists, prank, tq, ctr, total;
    uint32_t cap_slots;
};

static VcLayout vc_layout(int64_t lmax)
{
    VcLayout l;
    if (lmax < 1) lmax = 1;
    const size_t lp = (size_t)lmax + 64;
    l.cap_slots = (uint32_t)(lmax + lmax / 2 + 64);
    size_t o = 0;
    l.tab = o;   o += align_up((size_t)l.cap_slots * sizeof(VcEntry));
    l.pw = o;    o += align_up(lp * 4);
    l.pkey = o;  o += align_up(lp * 4);
    l.cw = o;    o += align_up(lp * 4);
    l.lists = o; o += align_up((3 * lp) * 4);   // sum over crowded voxels (>= 2 points) of count + 3 header words
    l.prank = o; o += align_up(lp * 4);
    l.tq = o;    o += align_up((lp / 2 + 64) * 4);   // crowded voxels hold >= 2 points
    l.ctr = o;   o += 256;
    l.total = o;
    return l;
}

struct VcArgs {
    const float *pts; int nfeat; const int64_t *offs; int64_t nframes;
    VoxCfg cfg;
    float *out_points; int64_t *out_mask; int64_t *out_mapping; int32_t *out_npoints; int64_t *out_coords; int64_t *counts;
    float *voxels; uint8_t *pmask;
    char *ws; VcLayout lay;
    unsigned long long *vstate, *kstate;   // sparse: cross-frame look-back words, one per frame (zeroed before the launch)
};

__device__ __forceinline__ VcEntry vc_load_entry(const VcEntry *p)
{
    uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p));   // L2 only: other CTAs of the cluster write these
    VcEntry e;
    e.kf = ((unsigned long long)v.y << 32) | v.x;
    e.count = v.z; e.aux = v.w;
    return e;
}

__device__ __forceinline__ float4 vc_load_point(const float *pts, int nfeat, int64_t i)
{
    if (nfeat == 4) return __ldg(reinterpret_cast<const float4 *>(pts) + i);
    const float *q = pts + i * nfeat;
    return make_float4(q[0], q[1], q[2], 0.f);
}

// Cross-frame decoupled look-back (sparse path): the outputs of all frames are packed back to back, so frame f
// starts at the sum of the row counts of frames < f.  Frames are processed in index order by co-resident
// clusters, every frame publishes its own count as soon as it knows it (flag 1) and its inclusive prefix once
// it has looked back (flag 2); a frame only ever waits for lower-numbered frames, which are running or done.
__device__ __forceinline__ unsigned long long vc_lookback(unsigned long long *state, int64_t f, unsigned long long mine, bool publish)
{
    constexpr unsigned long long VAL = (1ull << 62) - 1;
    if (publish && f > 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(state + f), "l"((1ull << 62) | mine) : "memory");
    unsigned long long excl = 0;
    for (int64_t j = f - 1; j >= 0; j--) {
        unsigned long long v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(state + j) : "memory");
            if (v >> 62) break;
            __nanosleep(64);
        }
        excl += v & VAL;
        if ((v >> 62) == 2) break;
    }
    if (publish) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(state + f), "l"((2ull << 62) | (excl + mine)) : "memory");
    return excl;
}

// exchange the per-warp totals of every CTA of the cluster through distributed shared memory and turn them
// into exclusive bases: base[v] = sum of totals of warps < v, base[W] = grand total
__device__ __forceinline__ void vc_exchange(cg::cluster_group &cluster, uint32_t *mytot, uint32_t *wt, uint32_t *base, unsigned csize, unsigned crank)
{
    const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    __syncthreads();   // mytot[] complete
    for (unsigned t = tid; t < csize * VC_WARPS; t += VC_THREADS) {
        uint32_t *remote = cluster.map_shared_rank(wt, t / VC_WARPS);
        remote[crank * VC_WARPS + (t % VC_WARPS)] = mytot[t % VC_WARPS];
    }
    cluster.sync();
    if (w == 0) {      // lane l owns totals [l*csize, (l+1)*csize): W = 32*csize values in warp order
        uint32_t s = 0;
        for (unsigned j = 0; j < csize; j++) s += wt[lane * csize + j];
        uint32_t inc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += t; }
        uint32_t run = inc - s;
        for (unsigned j = 0; j < csize; j++) { base[lane * csize + j] = run; run += wt[lane * csize + j]; }
        if (lane == 31) base[csize * VC_WARPS] = run;
    }
    __syncthreads();
}

// per-launch constants in the 32-bit form the kernel computes with
struct VcDev {
    float size[3], lo[3];
    int vlo[3];
    uint32_t ext[3];
    uint32_t sh_x, sh_y;     // cell key = cx << sh_x | cy << sh_y | cz (bit fields: decoding is two shifts and two masks)
    long long cadd[3];       // coords_out = c + vlo - offset
};

#ifndef D3D_VC_U
#define D3D_VC_U 2
#endif
constexpr int VC_U = D3D_VC_U;     // points per thread in flight in the streaming phases: independent loads are issued back to back
constexpr int VC_QCAP = 128;       // per-warp ring of points waiting for their next hash probe
constexpr uint32_t VC_NOKEY = 0xffffffffu;   // cell keys use at most 31 bits

template <bool DENSE>
__device__ __forceinline__ bool vc_cell(const VcDev &c, const float4 &p, uint32_t *key)
{
    float v0, v1, v2;
    if (DENSE) {
        v0 = __fdiv_rn(__fsub_rn(p.x, c.lo[0]), c.size[0]);
        v1 = __fdiv_rn(__fsub_rn(p.y, c.lo[1]), c.size[1]);
        v2 = __fdiv_rn(__fsub_rn(p.z, c.lo[2]), c.size[2]);
    } else {
        v0 = floorf(__fdiv_rn(p.x, c.size[0]));
        v1 = floorf(__fdiv_rn(p.y, c.size[1]));
        v2 = floorf(__fdiv_rn(p.z, c.size[2]));
    }
    const uint32_t c0 = (uint32_t)((int)v0 - c.vlo[0]), c1 = (uint32_t)((int)v1 - c.vlo[1]), c2 = (uint32_t)((int)v2 - c.vlo[2]);
    *key = (c0 << c.sh_x) | (c1 << c.sh_y) | c2;
    return v0 == v0 && v1 == v1 && v2 == v2 && c0 < c.ext[0] && c1 < c.ext[1] && c2 < c.ext[2];
}

#ifdef D3D_VC_TIMING   // tuning build only: per-phase time of a few frames, printed once per frame
#define VC_TICK(slot) do { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tk_[slot] = (float)(t_ - t0_) * 1e-3f; t0_ = t_; } while (0)
#define VC_TICK_INIT float tk_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; unsigned long long t0_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0_));
#define VC_TICK_PRINT do { if (ct == 0 && (cid == 0 || cid == ncl - 1) && f < 2 * (int64_t)ncl) \
    printf("frame %3d cl %2u  clear %6.1f insert %6.1f flags %6.1f join %6.1f big %6.1f ids %6.1f compact %6.1f us\n", (int)f, cid, tk_[0], tk_[1], tk_[2], tk_[3], tk_[4], tk_[5], tk_[6]); } while (0)
#else
#define VC_TICK(slot)
#define VC_TICK_INIT
#define VC_TICK_PRINT
#endif

template <bool DENSE>
__global__ void __launch_bounds__(VC_THREADS, 1) vox_cluster_kernel(const VcArgs a, const VcDev dv)
{
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned csize = cluster.num_blocks(), crank = cluster.block_rank();
    const unsigned ncl = gridDim.x / csize, cid = blockIdx.x / csize;
    const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned W = csize * VC_WARPS, g = crank * VC_WARPS + w;     // warps per cluster, index of this warp
    const unsigned CT = csize * VC_THREADS, ct = crank * VC_THREADS + tid;
    const unsigned ltmask = lanemask_lt();

    __shared__ uint32_t mytot[VC_WARPS];
    __shared__ uint32_t wt1[VC_MAX_CSIZE * VC_WARPS], wt2[VC_MAX_CSIZE * VC_WARPS];
    __shared__ uint32_t vbase[VC_MAX_CSIZE * VC_WARPS + 1], pbase[VC_MAX_CSIZE * VC_WARPS + 1];
    __shared__ unsigned long long frow[2];   // first voxel row / first kept-point row of this frame in the packed outputs
    __shared__ uint32_t q_key[VC_WARPS][VC_QCAP], q_pt[VC_WARPS][VC_QCAP];   // insert queue: cell key, (point index << 5 | probe number)

    char *slice = a.ws + (size_t)cid * a.lay.total;
    VcEntry *tab = reinterpret_cast<VcEntry *>(slice + a.lay.tab);
    uint32_t *pw = reinterpret_cast<uint32_t *>(slice + a.lay.pw);
    uint32_t *pkey = reinterpret_cast<uint32_t *>(slice + a.lay.pkey);
    uint32_t *cw = reinterpret_cast<uint32_t *>(slice + a.lay.cw);
    uint32_t *lists = reinterpret_cast<uint32_t *>(slice + a.lay.lists);
    uint32_t *prank = reinterpret_cast<uint32_t *>(slice + a.lay.prank);
    uint32_t *tq = reinterpret_cast<uint32_t *>(slice + a.lay.tq);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(slice + a.lay.ctr);   // [1] list cursor, [2] crowded voxels queued for ranking

    const VoxCfg &cfg = a.cfg;
    const uint32_t K = cfg.max_points > 0 ? (uint32_t)cfg.max_points : 0u;
    const bool trim = DENSE || cfg.pfilter == D3D_PF_TRIM;
    const bool dropall = !DENSE && trim && K == 0;                       // sparse TRIM with max_points 0 keeps no point
    const uint32_t cthr = DENSE ? 1u : ((trim && K > 0) ? K : VC_NONE);  // voxels with more points than this need ranks
    const uint32_t vcap = (DENSE || cfg.vfilter != D3D_VF_NONE) ? (cfg.max_voxels > 0 ? (uint32_t)cfg.max_voxels : 0u) : VC_NONE;
    const int nfeat = a.nfeat;

    for (int64_t f = cid; f < a.nframes; f += ncl) {
        const int64_t b = a.offs[f];
        const uint32_t L = (uint32_t)(a.offs[f + 1] - b);
        const uint32_t nslots = L + (L >> 1) + 64;
        const uint32_t nit = (L + W * 32 - 1) / (W * 32);   // 32-point rounds per warp
        const uint32_t wbeg = g * nit * 32;                 // this warp owns points [wbeg, wbeg + nit*32)

        VC_TICK_INIT
        // ---- P0: clear the table; cell key of every point (the reference's fp32 arithmetic), streamed with VC_U loads in flight
        for (uint32_t s = ct; s < nslots; s += CT)
            *reinterpret_cast<uint4 *>(tab + s) = make_uint4(0xffffffffu, 0xffffffffu, 0u, VC_NONE);
        if (ct < 16) ctr[ct] = 0;
        for (uint32_t k0 = 0; k0 < nit; k0 += VC_U) {
            float4 p[VC_U];
            uint32_t idx[VC_U];
            bool in[VC_U];
#pragma unroll
            for (int u = 0; u < VC_U; u++) {
                idx[u] = wbeg + (k0 + u) * 32 + lane;
                in[u] = k0 + u < nit && idx[u] < L;
                p[u] = in[u] ? vc_load_point(a.pts, nfeat, b + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < VC_U; u++) {
                uint32_t key;
                const bool ok = vc_cell<DENSE>(dv, p[u], &key);
                if (in[u]) pkey[idx[u]] = ok ? key : VC_NOKEY;
            }
        }
        cluster.sync();
        VC_TICK(0);

        // ---- P1: hash insert (smallest point index wins), points that join an existing voxel count themselves.
        // One probe = one atomic round trip to L2, and a lane may need several.  Instead of letting 31 lanes wait for
        // the unluckiest one, points wait in a per-warp shared-memory ring between probes and every CAS instruction
        // is issued for 32 queued points (the same compaction the IoU kernel uses for its candidate pairs).
        {
            uint32_t *qk = q_key[w], *qp = q_pt[w];
            uint32_t head = 0, tail = 0;   // warp-uniform
            for (uint32_t k = 0; k <= nit; k++) {
                if (k < nit) {
                    const uint32_t i = wbeg + k * 32 + lane;
                    const uint32_t key = i < L ? pkey[i] : VC_NOKEY;
                    const bool ok = key != VC_NOKEY;
                    const unsigned bal = __ballot_sync(0xffffffffu, ok);
                    if (ok) { const uint32_t t = (tail + __popc(bal & ltmask)) & (VC_QCAP - 1); qk[t] = key; qp[t] = i << 5; }
                    else if (i < L) pw[i] = VC_NONE;
                    tail += __popc(bal);
                }
                while (tail - head >= 32u || (k == nit && tail != head)) {
                    __syncwarp();
                    const uint32_t cnt = min(tail - head, 32u);
                    const bool act = lane < cnt;
                    const uint32_t t = (head + lane) & (VC_QCAP - 1);
                    const uint32_t key = act ? qk[t] : 0u, pt = act ? qp[t] : 0u;
                    head += cnt;
                    const uint32_t i = pt >> 5, probe = pt & 31u;
                    uint32_t s = __umulhi(key * 0x9E3779B1u, nslots) + probe;
                    if (s >= nslots) s -= nslots;
                    bool again = false;
                    if (act) {
                        const unsigned long long mine = ((unsigned long long)key << 32) | i;
                        unsigned long long c = atomicCAS(&tab[s].kf, VC_EMPTY, mine);
                        if (probe == 31u) {   // out of probe-counter bits (never seen in practice): finish this point in place
                            while (c != VC_EMPTY && (uint32_t)(c >> 32) != key) {
                                if (++s == nslots) s = 0;
                                c = atomicCAS(&tab[s].kf, VC_EMPTY, mine);
                            }
                        }
                        if (c == VC_EMPTY) {
                            pw[i] = s;                                              // claimed an empty slot
                        } else if ((uint32_t)(c >> 32) == key) {
                            if (i < (uint32_t)c) atomicMin(&tab[s].kf, mine);       // results unused: REDs
                            atomicAdd(&tab[s].count, 1u);
                            pw[i] = s;
                        } else {
                            again = true;                                           // another voxel's slot: probe the next one later
                        }
                    }
                    __syncwarp();
                    const unsigned bal = __ballot_sync(0xffffffffu, again);
                    if (again) { const uint32_t t2 = (tail + __popc(bal & ltmask)) & (VC_QCAP - 1); qk[t2] = key; qp[t2] = pt + 1; }
                    tail += __popc(bal);
                }
            }
        }
        cluster.sync();
        VC_TICK(1);

        // ---- P3: first-of-voxel flags -> provisional voxel ids (rank inside the warp's chunk).  The first point of a
        // crowded voxel reserves the voxel's list record and queues the voxel for ranking; the points of crowded
        // voxels are compacted into cw[] so that the next phase touches only them.
        uint32_t ncw = 0;   // crowded points of this warp's chunk (warp-uniform, lives across the barriers)
        {
            uint32_t carry = 0;
            for (uint32_t k0 = 0; k0 < nit; k0 += VC_U) {
                uint32_t idx[VC_U], s[VC_U];
                VcEntry e[VC_U];
#pragma unroll
                for (int u = 0; u < VC_U; u++) {
                    idx[u] = wbeg + (k0 + u) * 32 + lane;
                    s[u] = (k0 + u < nit && idx[u] < L) ? pw[idx[u]] : VC_NONE;
                }
#pragma unroll
                for (int u = 0; u < VC_U; u++)
                    if (s[u] != VC_NONE) e[u] = vc_load_entry(tab + s[u]);
#pragma unroll
                for (int u = 0; u < VC_U; u++) {
                    bool f1 = false, first = false, crowded = false, single = false;
                    uint32_t total = 0;
                    if (s[u] != VC_NONE) {
                        total = e[u].count + 1;
                        single = e[u].count == 0;
                        crowded = total > cthr;
                        first = (uint32_t)e[u].kf == idx[u];
                        f1 = first && (DENSE || (long long)total >= (long long)cfg.min_points);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, f1), balc = __ballot_sync(0xffffffffu, crowded);
                    const uint32_t prov = f1 ? ((g << VC_LOCAL_BITS) | (carry + __popc(bal & ltmask))) : VC_NONE;
                    if (s[u] != VC_NONE) {
                        if (crowded) {
                            cw[wbeg + ncw + __popc(balc & ltmask)] = idx[u];
                            if (first) {
                                const uint32_t base = atomicAdd(&ctr[1], total + VC_LIST_HDR);
                                lists[base] = prov; lists[base + 1] = 0; lists[base + 2] = 0;
                                tab[s[u]].aux = base;
                                tq[atomicAdd(&ctr[2], 1u)] = s[u];
                            }
                        } else if (single) {
                            pw[idx[u]] = f1 ? (VC_SINGLE | prov) : VC_NONE;     // everything later phases need is in this word
                        } else if (f1) {
                            tab[s[u]].aux = prov;
                        }
                    }
                    carry += __popc(bal);
                    ncw += __popc(balc);
                }
            }
            if (lane == 0) mytot[w] = carry;
            vc_exchange(cluster, mytot, wt1, vbase, csize, crank);
            if (!DENSE) {
                if (tid == 0) frow[0] = vc_lookback(a.vstate, f, min(vbase[W], vcap), crank == 0);
                __syncthreads();
            }
        }
        VC_TICK(2);

        const uint32_t ntask = __ldcg(&ctr[2]);   // cluster-uniform
        if (ntask) {
            // ---- P3b: points of crowded voxels join their voxel's list (any order)
            for (uint32_t j0 = 0; j0 < ncw; j0 += 32) {
                const uint32_t j = j0 + lane;
                if (j < ncw) {
                    const uint32_t i = cw[wbeg + j];
                    const uint32_t base = __ldcg(&tab[pw[i]].aux);
                    lists[base + VC_LIST_HDR + atomicAdd(&lists[base + 2], 1u)] = i;
                }
            }
            cluster.sync();
            VC_TICK(3);

            // ---- P4: one warp per crowded voxel ranks its points by index (the list order is arbitrary, the ranks are not)
            for (uint32_t q = g; q < ntask; q += W) {
                const uint32_t s = __ldcg(&tq[q]);
                const VcEntry e = vc_load_entry(tab + s);
                const uint32_t *lst = lists + e.aux + VC_LIST_HDR;
                const uint32_t total = e.count + 1;
                if (total <= 32u) {   // one index per lane: rank = number of smaller indices
                    const uint32_t x = lane < total ? __ldcg(lst + lane) : VC_NONE;
                    uint32_t r = 0;
                    for (uint32_t l = 0; l < total; l++) r += __shfl_sync(0xffffffffu, x, l) < x ? 1u : 0u;
                    if (lane < total) prank[x] = r;
                } else {              // extract the K smallest indices, one minimum per sweep; the rest is dropped by the bound
                    uint32_t lb = 0;
                    for (uint32_t r = 0; r < K; r++) {
                        uint32_t m = VC_NONE;
                        for (uint32_t j = lane; j < total; j += 32) {
                            const uint32_t x = __ldcg(lst + j);
                            if (x >= lb && x < m) m = x;
                        }
#pragma unroll
                        for (int d = 16; d; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
                        if (m == VC_NONE) break;
                        if (lane == 0) prank[m] = r;
                        lb = m + 1;
                    }
                    if (lane == 0) lists[e.aux + 1] = lb;
                }
            }
            cluster.sync();
            VC_TICK(4);
        }

        // ---- P5: final voxel ids, per-voxel outputs, point keep decisions
        {
            uint32_t carry = 0;
            const uint32_t mask_y = (1u << (dv.sh_x - dv.sh_y)) - 1, mask_z = (1u << dv.sh_y) - 1;
            for (uint32_t k0 = 0; k0 < nit; k0 += VC_U) {
                uint32_t idx[VC_U], s[VC_U], prov[VC_U], ckey[VC_U];
                VcEntry e[VC_U];
                bool single[VC_U];
#pragma unroll
                for (int u = 0; u < VC_U; u++) {
                    idx[u] = wbeg + (k0 + u) * 32 + lane;
                    s[u] = (k0 + u < nit && idx[u] < L) ? pw[idx[u]] : VC_NONE;
                    single[u] = s[u] != VC_NONE && (s[u] & VC_SINGLE);
                }
#pragma unroll
                for (int u = 0; u < VC_U; u++) {
                    if (single[u]) ckey[u] = pkey[idx[u]];                      // coalesced: the voxel's row is written from here
                    else if (s[u] != VC_NONE) e[u] = vc_load_entry(tab + s[u]);
                }
#pragma unroll
                for (int u = 0; u < VC_U; u++) {
                    if (single[u]) prov[u] = s[u] & ~VC_SINGLE;
                    else if (s[u] != VC_NONE) prov[u] = e[u].count + 1 > cthr ? __ldcg(&lists[e[u].aux]) : e[u].aux;
                }
#pragma unroll
                for (int u = 0; u < VC_U; u++) {
                    const uint32_t i = idx[u];
                    bool keep = false;
                    uint32_t nid = VC_NONE, rank = 0;
                    if (s[u] != VC_NONE) {
                        const uint32_t c = single[u] ? 1u : e[u].count + 1;
                        const bool crowded = !single[u] && c > cthr;
                        const bool first = single[u] || (uint32_t)e[u].kf == i;
                        if (prov[u] != VC_NONE) {
                            nid = vbase[prov[u] >> VC_LOCAL_BITS] + (prov[u] & ((1u << VC_LOCAL_BITS) - 1));
                            if (nid >= vcap) nid = VC_NONE;
                        }
                        keep = nid != VC_NONE;
                        if (keep && crowded) {
                            if (c > 32u) keep = i < __ldcg(&lists[e[u].aux + 1]);   // big voxel: only the K smallest indices have a rank
                            if (keep) { rank = __ldcg(&prank[i]); keep = rank < K; }
                        }
                        if (DENSE) keep = keep && rank < K;
                        if (dropall) keep = false;
                        if (nid != VC_NONE && first) {   // first point of a kept voxel writes the voxel's row
                            const uint32_t key = single[u] ? ckey[u] : (uint32_t)(e[u].kf >> 32);
                            const int64_t o = DENSE ? f * (int64_t)cfg.max_voxels + nid : (int64_t)frow[0] + nid;
                            long long *co = reinterpret_cast<long long *>(a.out_coords) + o * 3;
                            __stcs(co + 0, (long long)(key >> dv.sh_x) + dv.cadd[0]);
                            __stcs(co + 1, (long long)((key >> dv.sh_y) & mask_y) + dv.cadd[1]);
                            __stcs(co + 2, (long long)(key & mask_z) + dv.cadd[2]);
                            __stcs(a.out_npoints + o, (!DENSE && cfg.pfilter == D3D_PF_TRIM && c > K) ? (int32_t)K : (int32_t)c);
                        }
                        if (DENSE && keep) {
                            const int64_t slot = (f * (int64_t)cfg.max_voxels + nid) * (int64_t)K + rank;
                            if (nfeat == 4) __stcs(reinterpret_cast<float4 *>(a.voxels) + slot, vc_load_point(a.pts, 4, b + i));
                            else for (int q = 0; q < nfeat; q++) a.voxels[slot * nfeat + q] = a.pts[(b + i) * nfeat + q];
                            a.pmask[slot] = 1;
                        }
                    }
                    if (!DENSE) {
                        if (k0 + u < nit && i < L) pw[i] = keep ? nid : VC_NONE;
                        carry += __popc(__ballot_sync(0xffffffffu, keep));
                    }
                }
            }
            if (DENSE) {
                if (ct == 0) { a.counts[2 * f] = 0; a.counts[2 * f + 1] = (long long)min(vbase[W], vcap); }
                cluster.sync();   // the next frame clears the table other CTAs may still be reading
                continue;
            }
            if (lane == 0) mytot[w] = carry;
            vc_exchange(cluster, mytot, wt2, pbase, csize, crank);
            if (tid == 0) frow[1] = vc_lookback(a.kstate, f, pbase[W], crank == 0);
            __syncthreads();
        }
        VC_TICK(5);

        // ---- P6: compaction of the kept points, in input order
        {
            int64_t run = (int64_t)frow[1] + pbase[g];
            for (uint32_t k0 = 0; k0 < nit; k0 += VC_U) {
                uint32_t idx[VC_U], nid[VC_U];
                float4 p[VC_U];
#pragma unroll
                for (int u = 0; u < VC_U; u++) {
                    idx[u] = wbeg + (k0 + u) * 32 + lane;
                    nid[u] = (k0 + u < nit && idx[u] < L) ? pw[idx[u]] : VC_NONE;
                }
                if (nfeat == 4) {
#pragma unroll
                    for (int u = 0; u < VC_U; u++)
                        if (nid[u] != VC_NONE) p[u] = vc_load_point(a.pts, 4, b + idx[u]);
                }
#pragma unroll
                for (int u = 0; u < VC_U; u++) {
                    const bool keep = nid[u] != VC_NONE;
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const int64_t o = run + __popc(bal & ltmask);
                        if (nfeat == 4) __stcs(reinterpret_cast<float4 *>(a.out_points) + o, p[u]);
                        else for (int q = 0; q < nfeat; q++) a.out_points[o * nfeat + q] = a.pts[(b + idx[u]) * nfeat + q];
                        __stcs(reinterpret_cast<long long *>(a.out_mask) + o, (long long)idx[u]);
                        __stcs(reinterpret_cast<long long *>(a.out_mapping) + o, (long long)nid[u]);
                    }
                    run += __popc(bal);
                }
            }
            if (ct == 0) {   // frame_rows[f] = {first kept-point row, first voxel row}; the last frame also writes the totals
                a.counts[2 * f] = (long long)frow[1]; a.counts[2 * f + 1] = (long long)frow[0];
                if (f == a.nframes - 1) { a.counts[2 * f + 2] = (long long)(frow[1] + pbase[W]); a.counts[2 * f + 3] = (long long)(frow[0] + min(vbase[W], vcap)); }
            }
        }
        VC_TICK(6); VC_TICK_PRINT;
    }
}

// ------------------------------------------------------------------ host side
static int bits_for(long long ext)   // bits needed for coordinates 0 .. ext-1
{
    int b = 0;
    while ((1ll << b) < ext) b++;
    return b;
}

static bool vc_make_dev(const VoxCfg &cfg, VcDev *d)
{
    int bits[3];
    for (int k = 0; k < 3; k++) {
        if (cfg.ext[k] <= 0 || cfg.ext[k] > (1ll << 30)) return false;
        if (cfg.vlo[k] <= -(1ll << 30) || cfg.vlo[k] >= (1ll << 30)) return false;   // 32-bit cell arithmetic in vc_cell
        bits[k] = bits_for(cfg.ext[k]);
        d->size[k] = cfg.size[k]; d->lo[k] = cfg.lo[k];
        d->vlo[k] = (int)cfg.vlo[k]; d->ext[k] = (uint32_t)cfg.ext[k];
        d->cadd[k] = cfg.vlo[k] - cfg.offset[k];
    }
    if (bits[2] < 1) bits[2] = 1;   // keep the masks well defined for single-cell extents
    if (bits[1] < 1) bits[1] = 1;
    if (bits[0] + bits[1] + bits[2] > 31) return false;               // cell keys use at most 31 bits (VC_NOKEY is all ones)
    d->sh_y = (uint32_t)bits[2];
    d->sh_x = (uint32_t)(bits[2] + bits[1]);
    return true;
}

bool vox_cluster_supported(const VoxCfg &cfg, int64_t total, int64_t nframes, int64_t max_frame_points)
{
    (void)total; (void)nframes;
    VcDev d;
    if (!vc_make_dev(cfg, &d)) return false;                           // bit-field cell keys, 32-bit cell arithmetic
    if (max_frame_points >= (1ll << 27)) return false;                 // provisional ids: 9 + 20 bits
    if (cfg.ext[0] <= 0 || cfg.ext[1] <= 0 || cfg.ext[2] <= 0) return false;
    if (!cfg.dense && cfg.vfilter == D3D_VF_DESCENDING) return false;  // needs a sort of the voxels
    if (cfg.dense && cfg.reduction != D3D_RED_NONE) return false;      // sequential float sums need sorted segments
    return true;
}

size_t vox_cluster_ws_bytes(int64_t total, int64_t nframes, int64_t max_frame_points)
{
    if (max_frame_points <= 0 || max_frame_points > total) max_frame_points = total;
    int64_t ncl = nframes < VC_MAX_CLUSTERS ? nframes : VC_MAX_CLUSTERS;
    if (ncl < 1) ncl = 1;
    return vc_layout(max_frame_points).total * (size_t)ncl + 256 + align_up((size_t)(nframes > 0 ? nframes : 1) * 16);
}

// Cluster shape: CTAs of one cluster must sit in one GPC, and a B200's GPCs do not all expose a multiple of 8
// SMs, so 8-CTA clusters leave SMs idle (15 clusters = 120 of 148 SMs on the boxes measured).  Pick the size
// (<= 8, portable) that occupies the most SMs; the kernel is written for any cluster size.
struct VcShape { int csize, max_clusters; };

template <bool DENSE>
static int vc_shape(VcShape *out)
{
    static VcShape cached = {0, 0};
    if (cached.csize == 0) {
        auto kern = vox_cluster_kernel<DENSE>;
        VcShape best = {0, 0};
        int forced = 0;
        if (const char *e = getenv("D3D_B200_VOX_CLUSTER")) forced = atoi(e);   // tuning override
        for (int cs = 8; cs >= 4; cs -= 2) {   // 8, 6, 4: smaller clusters put more frames in flight than L2 holds
            if (forced) cs = forced;
            cudaLaunchConfig_t lc = {};
            lc.blockDim = dim3(VC_THREADS, 1, 1);
            lc.gridDim = dim3(cs, 1, 1);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &lc) != cudaSuccess) { cudaGetLastError(); continue; }
            if (n * cs > best.csize * best.max_clusters) best = {cs, n};
            if (forced) break;
        }
        if (best.csize == 0) return D3D_ERR_CUDA;
        cached = best;
    }
    *out = cached;
    return D3D_OK;
}

template <bool DENSE>
static int vc_launch(VcArgs &args, int64_t nframes, size_t ws_bytes, cudaStream_t st)
{
    VcDev dv;
    if (!vc_make_dev(args.cfg, &dv)) return D3D_ERR_UNSUPPORTED;
    VcShape shape;
    int rc = vc_shape<DENSE>(&shape);
    if (rc) return rc;
    const int csize = shape.csize;
    auto kern = vox_cluster_kernel<DENSE>;
    cudaLaunchConfig_t lc = {};
    lc.blockDim = dim3(VC_THREADS, 1, 1);
    lc.dynamicSmemBytes = 0;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int64_t ncl = shape.max_clusters;
    if (const char *e = getenv("D3D_B200_VOX_MAXCL")) { int m = atoi(e); if (m > 0 && m < ncl) ncl = m; }   // tuning override
    if (ncl > nframes) ncl = nframes;
    if (ncl > VC_MAX_CLUSTERS) ncl = VC_MAX_CLUSTERS;
    const size_t state_bytes = align_up((size_t)nframes * 16);
    while (ncl > 1 && args.lay.total * (size_t)ncl + 256 + state_bytes > ws_bytes) ncl--;
    if (args.lay.total * (size_t)ncl + 256 + state_bytes > ws_bytes) return D3D_ERR_WORKSPACE;
    if (!DENSE) {   // look-back words live behind the cluster slices
        args.vstate = reinterpret_cast<unsigned long long *>(args.ws + align_up(args.lay.total * (size_t)ncl));
        args.kstate = args.vstate + nframes;
        D3D_CUDA_TRY(cudaMemsetAsync(args.vstate, 0, (size_t)nframes * 16, st));
    }
    lc.gridDim = dim3((unsigned)(ncl * csize), 1, 1);
    D3D_CUDA_TRY(cudaLaunchKernelEx(&lc, kern, args, dv));
    D3D_LAUNCHED();
    return D3D_OK;
}

int vox_cluster_sparse(const float *points, int64_t total, int nfeat, const int64_t *offs, int64_t nframes, int64_t max_frame_points, const VoxCfg &cfg,
                       float *out_points, int64_t *out_mask, int64_t *out_mapping, int32_t *out_npoints, int64_t *out_coords, int64_t *counts,
                       void *ws, size_t ws_bytes, cudaStream_t st)
{
    if (max_frame_points <= 0 || max_frame_points > total) max_frame_points = total;
    VcArgs a = {};
    a.pts = points; a.nfeat = nfeat; a.offs = offs; a.nframes = nframes; a.cfg = cfg;
    a.out_points = out_points; a.out_mask = out_mask; a.out_mapping = out_mapping; a.out_npoints = out_npoints; a.out_coords = out_coords;
    a.counts = counts; a.ws = (char *)ws; a.lay = vc_layout(max_frame_points);
    return vc_launch<false>(a, nframes, ws_bytes, st);
}

int vox_cluster_dense(const float *points, int64_t total, int nfeat, const int64_t *offs, int64_t nframes, int64_t max_frame_points, const VoxCfg &cfg,
                      float *voxels, int64_t *coords, uint8_t *pmask, int32_t *npoints, int64_t *counts, void *ws, size_t ws_bytes, cudaStream_t st)
{
    if (max_frame_points <= 0 || max_frame_points > total) max_frame_points = total;
    VcArgs a = {};
    a.pts = points; a.nfeat = nfeat; a.offs = offs; a.nframes = nframes; a.cfg = cfg;
    a.out_npoints = npoints; a.out_coords = coords; a.counts = counts; a.voxels = voxels; a.pmask = pmask;
    a.ws = (char *)ws; a.lay = vc_layout(max_frame_points);
    return vc_launch<true>(a, nframes, ws_bytes, st);
}

}  // namespace d3d
