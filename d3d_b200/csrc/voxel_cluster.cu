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

namespace cg = cooperative_groups;

namespace d3d {

constexpr int VC_THREADS = 1024;
constexpr int VC_WARPS = VC_THREADS / 32;
constexpr int VC_MAX_CSIZE = 16;
constexpr int VC_MAX_CLUSTERS = 40;            // frames in flight the workspace is sized for
constexpr uint32_t VC_NONE = 0xffffffffu;
constexpr unsigned long long VC_EMPTY = ~0ull;
constexpr uint32_t VC_SMALL = 32;              // crowded voxels up to this size are ranked by the points themselves
constexpr int VC_LOCAL_BITS = 20;              // provisional voxel id = (warp << 20) | rank inside the warp's chunk

struct __align__(16) VcEntry {
    unsigned long long kf;   // (cell key << 32) | smallest frame-local index of a point in the cell
    uint32_t count;          // points in the cell
    uint32_t aux;            // crowded voxel: start of its list; otherwise the provisional voxel id
};

struct VcLayout {            // byte offsets inside one cluster's workspace slice
    size_t tab, pw, pos, lists, prank, cq, bigq, ctr, total;
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
    l.pos = o;   o += align_up(lp * 4);
    l.lists = o; o += align_up((2 * lp) * 4);
    l.prank = o; o += align_up(lp * 4);
    l.cq = o;    o += align_up((lp / 2 + 64) * 4);
    l.bigq = o;  o += align_up((lp / 32 + 64) * 4);
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

template <bool DENSE>
__global__ void __launch_bounds__(VC_THREADS, 1) vox_cluster_kernel(const VcArgs a)
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

    char *slice = a.ws + (size_t)cid * a.lay.total;
    VcEntry *tab = reinterpret_cast<VcEntry *>(slice + a.lay.tab);
    uint32_t *pw = reinterpret_cast<uint32_t *>(slice + a.lay.pw);
    uint32_t *pos = reinterpret_cast<uint32_t *>(slice + a.lay.pos);
    uint32_t *lists = reinterpret_cast<uint32_t *>(slice + a.lay.lists);
    uint32_t *prank = reinterpret_cast<uint32_t *>(slice + a.lay.prank);
    uint32_t *cq = reinterpret_cast<uint32_t *>(slice + a.lay.cq);
    uint32_t *bigq = reinterpret_cast<uint32_t *>(slice + a.lay.bigq);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(slice + a.lay.ctr);   // [0] crowded voxels, [1] list cursor, [2] big voxels

    const VoxCfg &cfg = a.cfg;
    const uint32_t K = cfg.max_points > 0 ? (uint32_t)cfg.max_points : 0u;
    const bool trim = DENSE || cfg.pfilter == D3D_PF_TRIM;
    const bool dropall = !DENSE && trim && K == 0;                       // sparse TRIM with max_points 0 keeps no point
    const uint32_t cthr = DENSE ? 1u : ((trim && K > 0) ? K : VC_NONE);  // voxels with more points than this need ranks
    const uint32_t vcap = (DENSE || cfg.vfilter != D3D_VF_NONE) ? (cfg.max_voxels > 0 ? (uint32_t)cfg.max_voxels : 0u) : VC_NONE;
    const uint32_t ext1 = (uint32_t)cfg.ext[1], ext2 = (uint32_t)cfg.ext[2];
    const int nfeat = a.nfeat;

    for (int64_t f = cid; f < a.nframes; f += ncl) {
        const int64_t b = a.offs[f];
        const uint32_t L = (uint32_t)(a.offs[f + 1] - b);
        const uint32_t nslots = L + (L >> 1) + 64;
        const uint32_t nit = (L + W * 32 - 1) / (W * 32);   // 32-point rounds per warp
        const uint32_t wbeg = g * nit * 32;                 // this warp owns points [wbeg, wbeg + nit*32)

        // ---- P0: clear the table
        for (uint32_t s = ct; s < nslots; s += CT)
            *reinterpret_cast<uint4 *>(tab + s) = make_uint4(0xffffffffu, 0xffffffffu, 0u, VC_NONE);
        if (ct < 16) ctr[ct] = 0;
        cluster.sync();

        // ---- P1: cell of every point, hash insert (smallest index wins), count
        for (uint32_t k = 0; k < nit; k++) {
            const uint32_t i = wbeg + k * 32 + lane;
            if (i < L) {
                const float4 p = vc_load_point(a.pts, nfeat, b + i);
                unsigned long long lin;
                uint32_t slot = VC_NONE;
                if (vox_cell(cfg, p.x, p.y, p.z, &lin)) {
                    const uint32_t key = (uint32_t)lin;
                    const unsigned long long mine = ((unsigned long long)key << 32) | i;
                    uint32_t s = __umulhi(key * 0x9E3779B1u, nslots);
                    for (;;) {
                        const unsigned long long cur = atomicCAS(&tab[s].kf, VC_EMPTY, mine);
                        if (cur == VC_EMPTY) break;
                        if ((uint32_t)(cur >> 32) == key) {
                            if (i < (uint32_t)cur) atomicMin(&tab[s].kf, mine);
                            break;
                        }
                        if (++s == nslots) s = 0;
                    }
                    const uint32_t old = atomicAdd(&tab[s].count, 1u);
                    pos[i] = old;
                    if (old == cthr) cq[atomicAdd(&ctr[0], 1u)] = s;   // exactly one point per crowded voxel sees this
                    slot = s;
                }
                pw[i] = slot;
            }
        }
        cluster.sync();

        // ---- P2: list storage for crowded voxels (counts are final now)
        const uint32_t ncq = __ldcg(&ctr[0]);
        if (ncq) {
            for (uint32_t q = ct; q < ncq; q += CT) {
                const uint32_t s = __ldcg(&cq[q]);
                const uint32_t c = __ldcg(&tab[s].count);
                const uint32_t base = atomicAdd(&ctr[1], c + 2);
                tab[s].aux = base;
                lists[base] = VC_NONE;      // provisional voxel id
                lists[base + 1] = 0;        // big voxels: keep points with index below this
                if (c > VC_SMALL) bigq[atomicAdd(&ctr[2], 1u)] = s;
            }
            cluster.sync();
        }

        // ---- P3: first-of-voxel flags -> provisional voxel ids (rank inside the warp's chunk); crowded points join their list
        {
            uint32_t carry = 0;
            for (uint32_t k = 0; k < nit; k++) {
                const uint32_t i = wbeg + k * 32 + lane;
                const uint32_t s = i < L ? pw[i] : VC_NONE;
                bool f1 = false, crowded = false;
                uint32_t aux = 0;
                if (s != VC_NONE) {
                    const VcEntry e = vc_load_entry(tab + s);
                    crowded = e.count > cthr;
                    aux = e.aux;
                    if (crowded) lists[aux + 2 + pos[i]] = i;
                    f1 = (uint32_t)e.kf == i && (DENSE || (long long)e.count >= (long long)cfg.min_points);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, f1);
                if (f1) {
                    const uint32_t prov = (g << VC_LOCAL_BITS) | (carry + __popc(bal & ltmask));
                    if (crowded) lists[aux] = prov; else tab[s].aux = prov;
                }
                carry += __popc(bal);
            }
            if (lane == 0) mytot[w] = carry;
            vc_exchange(cluster, mytot, wt1, vbase, csize, crank);
        }

        // ---- P4: big crowded voxels: one warp extracts the K smallest point indices, one minimum per sweep
        const uint32_t nbig = __ldcg(&ctr[2]);
        if (nbig) {
            for (uint32_t q = g; q < nbig; q += W) {
                const uint32_t s = __ldcg(&bigq[q]);
                const VcEntry e = vc_load_entry(tab + s);
                const uint32_t *lst = lists + e.aux + 2;
                uint32_t lb = 0;
                for (uint32_t r = 0; r < K; r++) {
                    uint32_t m = VC_NONE;
                    for (uint32_t j = lane; j < e.count; j += 32) {
                        const uint32_t x = __ldcg(lst + j);
                        if (x >= lb && x < m) m = x;
                    }
#pragma unroll
                    for (int d = 16; d; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
                    if (m == VC_NONE) break;
                    if (DENSE && lane == 0) prank[m] = r;
                    lb = m + 1;
                }
                if (lane == 0) lists[e.aux + 1] = lb;
            }
            cluster.sync();
        }

        // ---- P5: final voxel ids, per-voxel outputs, point keep decisions
        {
            uint32_t carry = 0;
            for (uint32_t k = 0; k < nit; k++) {
                const uint32_t i = wbeg + k * 32 + lane;
                const uint32_t s = i < L ? pw[i] : VC_NONE;
                bool keep = false;
                uint32_t nid = VC_NONE, rank = 0;
                if (s != VC_NONE) {
                    const VcEntry e = vc_load_entry(tab + s);
                    const uint32_t c = e.count;
                    const bool crowded = c > cthr;
                    const uint32_t prov = crowded ? __ldcg(&lists[e.aux]) : e.aux;
                    if (prov != VC_NONE) {
                        nid = vbase[prov >> VC_LOCAL_BITS] + (prov & ((1u << VC_LOCAL_BITS) - 1));
                        if (nid >= vcap) nid = VC_NONE;
                    }
                    keep = nid != VC_NONE;
                    if (keep && crowded) {
                        if (c <= VC_SMALL) {
                            const uint32_t *lst = lists + e.aux + 2;
                            for (uint32_t j = 0; j < c; j++) rank += __ldcg(lst + j) < i ? 1u : 0u;
                            keep = rank < K;
                        } else {
                            keep = i < __ldcg(&lists[e.aux + 1]);
                            if (DENSE && keep) rank = __ldcg(&prank[i]);
                        }
                    }
                    if (DENSE) keep = keep && rank < K;
                    if (dropall) keep = false;
                    if (nid != VC_NONE && (uint32_t)e.kf == i) {   // first point of a kept voxel writes the voxel's row
                        uint32_t key = (uint32_t)(e.kf >> 32);
                        const uint32_t cz = key % ext2; key /= ext2;
                        const uint32_t cy = key % ext1;
                        const uint32_t cx = key / ext1;
                        const int64_t o = DENSE ? f * (int64_t)cfg.max_voxels + nid : b + nid;
                        a.out_coords[o * 3 + 0] = (long long)cx + cfg.vlo[0] - cfg.offset[0];
                        a.out_coords[o * 3 + 1] = (long long)cy + cfg.vlo[1] - cfg.offset[1];
                        a.out_coords[o * 3 + 2] = (long long)cz + cfg.vlo[2] - cfg.offset[2];
                        a.out_npoints[o] = (!DENSE && cfg.pfilter == D3D_PF_TRIM && c > K) ? (int32_t)K : (int32_t)c;
                    }
                    if (DENSE && keep) {
                        const int64_t slot = (f * (int64_t)cfg.max_voxels + nid) * (int64_t)K + rank;
                        if (nfeat == 4) reinterpret_cast<float4 *>(a.voxels)[slot] = vc_load_point(a.pts, 4, b + i);
                        else for (int q = 0; q < nfeat; q++) a.voxels[slot * nfeat + q] = a.pts[(b + i) * nfeat + q];
                        a.pmask[slot] = 1;
                    }
                }
                if (!DENSE) {
                    if (i < L) pw[i] = keep ? nid : VC_NONE;
                    carry += __popc(__ballot_sync(0xffffffffu, keep));
                }
            }
            if (DENSE) {
                if (ct == 0) { a.counts[2 * f] = 0; a.counts[2 * f + 1] = (long long)min(vbase[W], vcap); }
                cluster.sync();   // the next frame clears the table other CTAs may still be reading
                continue;
            }
            if (lane == 0) mytot[w] = carry;
            vc_exchange(cluster, mytot, wt2, pbase, csize, crank);
        }

        // ---- P6: compaction of the kept points, in input order
        {
            uint32_t run = pbase[g];
            for (uint32_t k = 0; k < nit; k++) {
                const uint32_t i = wbeg + k * 32 + lane;
                const uint32_t nid = i < L ? pw[i] : VC_NONE;
                const bool keep = nid != VC_NONE;
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int64_t o = b + run + __popc(bal & ltmask);
                    if (nfeat == 4) reinterpret_cast<float4 *>(a.out_points)[o] = vc_load_point(a.pts, 4, b + i);
                    else for (int q = 0; q < nfeat; q++) a.out_points[o * nfeat + q] = a.pts[(b + i) * nfeat + q];
                    a.out_mask[o] = i;
                    a.out_mapping[o] = nid;
                }
                run += __popc(bal);
            }
            if (ct == 0) { a.counts[2 * f] = pbase[W]; a.counts[2 * f + 1] = (long long)min(vbase[W], vcap); }
        }
    }
}

// ------------------------------------------------------------------ host side
static int vc_cluster_size()
{
    return 8;   // portable maximum; one CTA per SM, 18 frames in flight on 148 SMs
}

bool vox_cluster_supported(const VoxCfg &cfg, int64_t total, int64_t nframes, int64_t max_frame_points)
{
    (void)total; (void)nframes;
    if (cfg.G >= (1ull << 32) - 1) return false;                       // 32-bit cell keys
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
    return vc_layout(max_frame_points).total * (size_t)ncl + 256;
}

template <bool DENSE>
static int vc_launch(VcArgs &args, int64_t nframes, size_t ws_bytes, cudaStream_t st)
{
    const int csize = vc_cluster_size();
    auto kern = vox_cluster_kernel<DENSE>;
    cudaLaunchConfig_t lc = {};
    lc.blockDim = dim3(VC_THREADS, 1, 1);
    lc.gridDim = dim3(csize, 1, 1);
    lc.dynamicSmemBytes = 0;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    static int max_clusters[2] = {0, 0};
    if (max_clusters[DENSE] == 0) {
        int n = 0;
        D3D_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kern, &lc));
        if (n < 1) n = 1;
        max_clusters[DENSE] = n;
    }
    int64_t ncl = max_clusters[DENSE];
    if (ncl > nframes) ncl = nframes;
    if (ncl > VC_MAX_CLUSTERS) ncl = VC_MAX_CLUSTERS;
    while (ncl > 1 && args.lay.total * (size_t)ncl + 256 > ws_bytes) ncl--;
    if (args.lay.total * (size_t)ncl + 256 > ws_bytes) return D3D_ERR_WORKSPACE;
    lc.gridDim = dim3((unsigned)(ncl * csize), 1, 1);
    D3D_CUDA_TRY(cudaLaunchKernelEx(&lc, kern, args));
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
