// voxel.cu -- deterministic voxelization of a batch of point-cloud frames (sparse+filter and dense).
//
// Replaces reference voxelize_sparse + voxelize_filter and voxelize_3d_dense
// (d3d/voxel/voxelize.cpp:288-484 and :45-199), which are single-threaded std::unordered_map loops;
// the reference has no CUDA voxelizer.  The reference's semantics are sequential ("voxel ids in order
// of first appearance", "first max_points points of a voxel in input order"), so the GPU formulation
// is built from order-preserving primitives only:
//   1. key pass (float4-vectorised when nfeat == 4): voxel coordinate with the reference's exact fp32
//      arithmetic (IEEE divide then floor / truncate), bounds test, 64-bit key = frame*G + linear cell;
//   2. our stable LSD radix sort of (key, point index) on the significant key bits (prims.cu):
//      a voxel's points become one segment, in input order;
//   3. segment heads -> scan -> per-segment start / first point; first-point flags scanned in POINT
//      order give every voxel its first-appearance rank inside its frame;
//   4. voxel filter (min_points, TRIM, DESCENDING) by a scan (or a second sort) in appearance order;
//   5. point filter (rank inside the segment < max_points), scan in point order, compaction.
// No atomics decide any output, so every tensor is bit-reproducible run to run and equal to the
// reference's: coords, points_mapping, points_mask, voxel_npoints, dense voxels/pmask/aggregates
// (MEAN sums a voxel's points sequentially in input order like voxelize.cpp:137-164).
#include "voxel.cuh"
#include "prims.cuh"

namespace d3d {

__device__ __forceinline__ int64_t frame_of(const int64_t *__restrict__ offs, int64_t nframes, int64_t i)
{
    int64_t lo = 0, hi = nframes;   // largest f with offs[f] <= i
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (offs[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(256) vox_key_kernel(const float *__restrict__ pts, int64_t total, int nfeat, const int64_t *__restrict__ offs,
                                                      int64_t nframes, VoxCfg cfg, uint64_t *__restrict__ keys, uint32_t *__restrict__ idx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float p[3];
    if (nfeat == 4) {   // 16-byte coalesced point load
        float4 v = __ldg(reinterpret_cast<const float4 *>(pts) + i);
        p[0] = v.x; p[1] = v.y; p[2] = v.z;
    } else {
        const float *q = pts + i * nfeat;
        p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
    }
    unsigned long long lin;
    const bool ok = vox_cell(cfg, p[0], p[1], p[2], &lin);
    uint64_t key = VOX_INVALID;
    if (ok) key = (uint64_t)frame_of(offs, nframes, i) * cfg.G + lin;
    keys[i] = key;
    idx[i] = (uint32_t)i;
}

// flag[p] = 1 at every segment head of the sorted key array, and at the terminator p == total
__global__ void __launch_bounds__(256) vox_head_kernel(const uint64_t *__restrict__ keys, int64_t total, uint32_t *__restrict__ flag)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > total) return;
    flag[p] = (p == 0 || p == total || keys[p] != keys[p - 1]) ? 1u : 0u;
}

// per head: segment start and first (lowest-index) point; mark that point in point order
__global__ void __launch_bounds__(256) vox_seg_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ sidx, const uint32_t *__restrict__ flag,
                                                      uint32_t *__restrict__ segid, int64_t total, uint32_t *__restrict__ seg_start,
                                                      uint32_t *__restrict__ seg_first, uint32_t *__restrict__ firstflag)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > total) return;
    uint32_t s = segid[p];
    // segid came from an EXCLUSIVE scan of the head flags: a non-head position has already counted the
    // head of its own segment, so its segment index is one less
    if (!flag[p]) { segid[p] = s - 1; return; }
    seg_start[s] = (uint32_t)p;
    if (p < total) {
        seg_first[s] = sidx[p];
        if (keys[p] != VOX_INVALID) firstflag[sidx[p]] = 1u;
    }
}

// per voxel segment: appearance rank a (global), inverse map, "passes the voxel filter" flag in appearance order
__global__ void __launch_bounds__(256) vox_pass_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ seg_first,
                                                       const uint32_t *__restrict__ app, const uint32_t *__restrict__ nseg_ptr, VoxCfg cfg, uint32_t *__restrict__ seg_app,
                                                       uint32_t *__restrict__ passflag)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (int64_t)nseg_ptr[0]) return;
    uint32_t p0 = seg_start[s];
    if (keys[p0] == VOX_INVALID) return;
    uint32_t cnt = seg_start[s + 1] - p0;
    uint32_t a = app[seg_first[s]];
    seg_app[s] = a;
    passflag[a] = (cfg.dense || (int64_t)cnt >= (int64_t)cfg.min_points) ? 1u : 0u;
}

// DESCENDING: key = (frame, count descending, appearance ascending) per voxel segment
__global__ void __launch_bounds__(256) vox_desc_key_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ seg_app,
                                                           const uint32_t *__restrict__ app, const int64_t *__restrict__ offs, const uint32_t *__restrict__ nseg_ptr, VoxCfg cfg,
                                                           uint64_t *__restrict__ dkeys, uint32_t *__restrict__ dvals)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (int64_t)nseg_ptr[0]) return;
    uint32_t p0 = seg_start[s];
    dvals[s] = (uint32_t)s;
    if (keys[p0] == VOX_INVALID) { dkeys[s] = 0x7fffffffffffffffull; return; }
    uint64_t f = keys[p0] / cfg.G;
    uint32_t cnt = seg_start[s + 1] - p0;
    uint32_t al = seg_app[s] - app[offs[f]];
    uint32_t inv = 0xfffffu - (cnt > 0xfffffu ? 0xfffffu : cnt);
    dkeys[s] = (f << 52) | ((uint64_t)inv << 32) | (uint64_t)al;
}

// position q in the DESCENDING order -> new id of that voxel
__global__ void __launch_bounds__(256) vox_desc_rank_kernel(const uint64_t *__restrict__ dkeys, const uint32_t *__restrict__ dvals, const uint32_t *__restrict__ app,
                                                            const int64_t *__restrict__ offs, const uint32_t *__restrict__ nseg_ptr, uint32_t *__restrict__ desc_rank)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)nseg_ptr[0]) return;
    uint64_t k = dkeys[q];
    if (k == 0x7fffffffffffffffull) return;
    uint64_t f = k >> 52;
    desc_rank[dvals[q]] = (uint32_t)q - app[offs[f]];   // frame f's voxels occupy [app_base(f), app_base(f+1)) in this order too
}

// per voxel segment: new id (or -1) and the per-voxel outputs
__global__ void __launch_bounds__(256) vox_voxel_out_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ seg_app,
                                                            const uint32_t *__restrict__ app, const uint32_t *__restrict__ prank, const uint32_t *__restrict__ passflag,
                                                            const uint32_t *__restrict__ desc_rank, const int64_t *__restrict__ offs, const int64_t *__restrict__ voff,
                                                            const uint32_t *__restrict__ nseg_ptr, VoxCfg cfg,
                                                            int32_t *__restrict__ newid, int32_t *__restrict__ out_npoints, int64_t *__restrict__ out_coords)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (int64_t)nseg_ptr[0]) return;
    uint32_t p0 = seg_start[s];
    uint64_t key = keys[p0];
    if (key == VOX_INVALID) { newid[s] = -1; return; }
    uint64_t f = key / cfg.G, lin = key % cfg.G;
    uint32_t cnt = seg_start[s + 1] - p0;
    uint32_t a = seg_app[s];
    uint32_t abase = app[offs[f]];
    int64_t nid;
    bool keep = passflag[a] != 0;
    if (cfg.dense) { nid = (int64_t)a - abase; keep = nid < cfg.max_voxels; }
    else if (cfg.vfilter == D3D_VF_DESCENDING) { nid = desc_rank[s]; keep = keep && nid < cfg.max_voxels; }
    else { nid = (int64_t)prank[a] - prank[abase]; keep = keep && (cfg.vfilter == D3D_VF_NONE || nid < cfg.max_voxels); }
    newid[s] = keep ? (int32_t)nid : -1;
    if (!keep) return;
    long long cz = (long long)(lin % (unsigned long long)cfg.ext[2]); lin /= (unsigned long long)cfg.ext[2];
    long long cy = (long long)(lin % (unsigned long long)cfg.ext[1]); lin /= (unsigned long long)cfg.ext[1];
    long long cx = (long long)lin;
    int64_t o = cfg.dense ? (int64_t)f * cfg.max_voxels + nid : voff[f] + nid;   // sparse outputs are packed across frames
    out_coords[o * 3 + 0] = cx + cfg.vlo[0] - cfg.offset[0];
    out_coords[o * 3 + 1] = cy + cfg.vlo[1] - cfg.offset[1];
    out_coords[o * 3 + 2] = cz + cfg.vlo[2] - cfg.offset[2];
    out_npoints[o] = (!cfg.dense && cfg.pfilter == D3D_PF_TRIM && (int64_t)cnt > cfg.max_points) ? cfg.max_points : (int32_t)cnt;
}

// sparse: per sorted position -> keep flag and new voxel id of the point, in POINT order
__global__ void __launch_bounds__(256) vox_point_flag_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ sidx, const uint32_t *__restrict__ segid,
                                                             const uint32_t *__restrict__ seg_start, const int32_t *__restrict__ newid, int64_t total, VoxCfg cfg,
                                                             uint32_t *__restrict__ keepflag, int32_t *__restrict__ pmap)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    uint32_t i = sidx[p];
    uint32_t keep = 0; int32_t nid = -1;
    if (keys[p] != VOX_INVALID) {
        uint32_t s = segid[p];
        nid = newid[s];
        uint32_t rank = (uint32_t)p - seg_start[s];
        keep = (nid >= 0 && (cfg.pfilter == D3D_PF_NONE || (int64_t)rank < cfg.max_points)) ? 1u : 0u;
    }
    keepflag[i] = keep;
    pmap[i] = nid;
}

__global__ void __launch_bounds__(256) vox_point_out_kernel(const float *__restrict__ pts, int nfeat, const uint32_t *__restrict__ keepflag, const uint32_t *__restrict__ ppos,
                                                            const int32_t *__restrict__ pmap, const int64_t *__restrict__ offs, int64_t nframes, int64_t total,
                                                            float *__restrict__ out_points, int64_t *__restrict__ out_mask, int64_t *__restrict__ out_mapping)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total || !keepflag[i]) return;
    int64_t f = frame_of(offs, nframes, i);
    int64_t o = (int64_t)ppos[i];   // global exclusive scan of the keep flags: rows are packed across frames
    if (nfeat == 4) reinterpret_cast<float4 *>(out_points)[o] = __ldg(reinterpret_cast<const float4 *>(pts) + i);
    else for (int k = 0; k < nfeat; k++) out_points[o * nfeat + k] = pts[i * nfeat + k];
    out_mask[o] = i - offs[f];
    out_mapping[o] = pmap[i];
}

// voff[f] = first voxel row of frame f in the packed outputs = sum over earlier frames of their (capped) voxel counts
__global__ void __launch_bounds__(1024) vox_voxel_rows_kernel(const uint32_t *__restrict__ app, const uint32_t *__restrict__ prank, const int64_t *__restrict__ offs,
                                                              int64_t nframes, VoxCfg cfg, int64_t *__restrict__ voff)
{
    __shared__ long long wsum[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < nframes; base += 1024) {
        const int64_t f = base + threadIdx.x;
        long long nv = 0;
        if (f < nframes) {
            nv = (long long)(prank[app[offs[f + 1]]] - prank[app[offs[f]]]);   // voxels passing min_points (and bounds)
            if (cfg.vfilter != D3D_VF_NONE && nv > cfg.max_voxels) nv = cfg.max_voxels;
        }
        long long inc = nv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { long long t = __shfl_up_sync(0xffffffffu, inc, d); if (lane_id() >= (unsigned)d) inc += t; }
        if (lane_id() == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            long long v = wsum[threadIdx.x], vi = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { long long t = __shfl_up_sync(0xffffffffu, vi, d); if (lane_id() >= (unsigned)d) vi += t; }
            wsum[threadIdx.x] = vi - v;
        }
        __syncthreads();
        const long long excl = carry_s + wsum[threadIdx.x >> 5] + inc - nv;
        if (f < nframes) voff[f] = excl;
        __syncthreads();
        if (threadIdx.x == 1023 || f == nframes - 1) { carry_s = excl + nv; if (f == nframes - 1) voff[nframes] = excl + nv; }
        __syncthreads();
    }
}

// frame_rows[f] = {first kept-point row, first voxel row} of frame f; entry nframes holds the totals
__global__ void vox_rows_sparse_kernel(const uint32_t *__restrict__ ppos, const int64_t *__restrict__ voff, const int64_t *__restrict__ offs, int64_t nframes,
                                       int64_t *__restrict__ rows)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f > nframes) return;
    rows[2 * f] = (int64_t)ppos[offs[f]];
    rows[2 * f + 1] = voff[f];
}

// dense: slot writer, one thread per sorted position
__global__ void __launch_bounds__(256) vox_dense_point_kernel(const float *__restrict__ pts, int nfeat, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ sidx,
                                                              const uint32_t *__restrict__ segid, const uint32_t *__restrict__ seg_start, const int32_t *__restrict__ newid,
                                                              int64_t total, VoxCfg cfg, float *__restrict__ voxels, uint8_t *__restrict__ pmask)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    uint64_t key = keys[p];
    if (key == VOX_INVALID) return;
    uint32_t s = segid[p];
    int32_t nid = newid[s];
    uint32_t rank = (uint32_t)p - seg_start[s];
    if (nid < 0 || (int64_t)rank >= cfg.max_points) return;
    int64_t f = (int64_t)(key / cfg.G);
    int64_t slot = ((int64_t)f * cfg.max_voxels + nid) * cfg.max_points + rank;
    uint32_t i = sidx[p];
    if (nfeat == 4) reinterpret_cast<float4 *>(voxels)[slot] = __ldg(reinterpret_cast<const float4 *>(pts) + i);
    else for (int k = 0; k < nfeat; k++) voxels[slot * nfeat + k] = pts[(int64_t)i * nfeat + k];
    pmask[slot] = 1;
}

// dense: aggregates over ALL points of a voxel, sequentially in input order (voxelize.cpp:137-164)
__global__ void __launch_bounds__(128) vox_dense_aggr_kernel(const float *__restrict__ pts, int nfeat, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ sidx,
                                                             const uint32_t *__restrict__ seg_start, const int32_t *__restrict__ newid, const uint32_t *__restrict__ nseg_ptr, VoxCfg cfg,
                                                             float *__restrict__ aggr)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t s = t / nfeat; int k = (int)(t % nfeat);
    if (s >= (int64_t)nseg_ptr[0]) return;
    int32_t nid = newid[s];
    if (nid < 0) return;
    uint32_t p0 = seg_start[s], p1 = seg_start[s + 1];
    int64_t f = (int64_t)(keys[p0] / cfg.G);
    float acc = cfg.reduction == D3D_RED_MEAN ? 0.0f : (cfg.reduction == D3D_RED_MAX ? -INFINITY : INFINITY);
    for (uint32_t p = p0; p < p1; p++) {
        float v = pts[(int64_t)sidx[p] * nfeat + k];
        if (cfg.reduction == D3D_RED_MEAN) acc = __fadd_rn(acc, v);
        else if (cfg.reduction == D3D_RED_MAX) acc = acc > v ? acc : v;   // std::max(a, p)
        else acc = v < acc ? v : acc;                                       // std::min(a, p)
    }
    if (cfg.reduction == D3D_RED_MEAN) acc = __fdiv_rn(acc, (float)(int)(p1 - p0));
    aggr[((int64_t)f * cfg.max_voxels + nid) * nfeat + k] = acc;
}

__global__ void vox_counts_dense_kernel(const uint32_t *__restrict__ app, const int64_t *__restrict__ offs, int64_t nframes, VoxCfg cfg, int64_t *__restrict__ counts)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    int64_t nv = (int64_t)(app[offs[f + 1]] - app[offs[f]]);
    if (nv > cfg.max_voxels) nv = cfg.max_voxels;
    counts[2 * f] = 0;
    counts[2 * f + 1] = nv;
}

static int key_bits_of(unsigned long long maxkey)
{
    int b = 0;
    while (maxkey) { b++; maxkey >>= 1; }
    return b < 1 ? 1 : b;
}

static size_t vox_ws_bytes(int64_t total, int64_t nframes)
{
    size_t n1 = (size_t)(total + 2);
    return align_up(n1 * 8) * 2 + align_up(n1 * 4) * 12 + radix_sort_workspace_bytes(total + 1) + scan_workspace_bytes(total + 2) +
           align_up((size_t)(nframes + 2) * 8) + 8192;
}

static int build_cfg(const d3d_voxel_params *P, int dense, int64_t nframes, VoxCfg *cfg)
{
    VoxCfg c;
    c.dense = dense;
    c.min_points = P->min_points; c.max_points = P->max_points; c.max_voxels = P->max_voxels;
    c.pfilter = P->max_points_filter; c.vfilter = P->max_voxels_filter; c.reduction = P->reduction;
    unsigned long long G = 1;
    for (int d = 0; d < 3; d++) {
        if (dense) {
            if (P->shape[d] <= 0) return D3D_ERR_INVALID_ARGUMENT;
            c.size[d] = (P->bound[2 * d + 1] - P->bound[2 * d]) / P->shape[d];   // float / int -> float, voxelize.cpp:85-87
            c.lo[d] = P->bound[2 * d];
            c.vlo[d] = 0; c.ext[d] = P->shape[d]; c.offset[d] = 0;
        } else {
            c.size[d] = P->size[d]; c.lo[d] = 0;
            c.vlo[d] = P->vlo[d]; c.ext[d] = P->vhi[d] - P->vlo[d]; c.offset[d] = P->offset[d];
            if (c.ext[d] < 0) c.ext[d] = 0;
        }
        unsigned long long e = (unsigned long long)(c.ext[d] > 0 ? c.ext[d] : 1);
        if (G > (1ull << 62) / e) return D3D_ERR_RANGE;
        G *= e;
    }
    if ((unsigned long long)(nframes > 0 ? nframes : 1) > (1ull << 62) / G) return D3D_ERR_RANGE;
    c.G = G;
    *cfg = c;
    return D3D_OK;
}

struct VoxBufs {
    uint64_t *keys, *dkeys;
    uint32_t *sidx, *flag, *segid, *seg_start, *seg_first, *app, *seg_app, *passflag, *prank, *keepflag, *ppos, *desc_rank;
    int32_t *newid, *pmap;
    void *sort_ws, *scan_ws;
    size_t sort_bytes;
    int64_t *voff;
};

static int vox_common(const float *points, int64_t total, int nfeat, const int64_t *offs, int64_t nframes, const VoxCfg &cfg, VoxBufs &B, void *ws, size_t ws_bytes,
                      cudaStream_t st)
{
    Arena a(ws, ws_bytes);
    size_t n1 = (size_t)(total + 2);
    B.keys = a.take<uint64_t>(n1); B.dkeys = a.take<uint64_t>(n1);
    B.sidx = a.take<uint32_t>(n1); B.flag = a.take<uint32_t>(n1); B.segid = a.take<uint32_t>(n1); B.seg_start = a.take<uint32_t>(n1);
    B.seg_first = a.take<uint32_t>(n1); B.app = a.take<uint32_t>(n1); B.seg_app = a.take<uint32_t>(n1); B.passflag = a.take<uint32_t>(n1);
    B.prank = a.take<uint32_t>(n1); B.keepflag = a.take<uint32_t>(n1); B.ppos = a.take<uint32_t>(n1); B.desc_rank = a.take<uint32_t>(n1);
    B.newid = (int32_t *)B.flag;      // flag is dead once segid exists
    B.pmap = (int32_t *)B.seg_first;  // seg_first is dead once seg_app exists
    B.sort_bytes = radix_sort_workspace_bytes(total + 1);
    B.sort_ws = a.take<char>(B.sort_bytes);
    B.scan_ws = a.take<char>(scan_workspace_bytes(total + 2));
    B.voff = a.take<int64_t>((size_t)nframes + 2);
    if (!a.ok()) return D3D_ERR_WORKSPACE;
    const unsigned gb = (unsigned)cdiv(total + 1, 256);
    int rc;
    vox_key_kernel<<<gb, 256, 0, st>>>(points, total, nfeat, offs, nframes, cfg, B.keys, B.sidx); D3D_LAUNCHED();
    // invalid keys are all-ones: sort on one bit more than the largest valid key so they end up last
    int bits = key_bits_of((unsigned long long)nframes * cfg.G) + 1;
    if (bits > 64) bits = 64;
    if ((rc = radix_sort_pairs_u64(B.keys, B.sidx, total, bits, B.sort_ws, B.sort_bytes, st))) return rc;
    vox_head_kernel<<<gb, 256, 0, st>>>(B.keys, total, B.flag); D3D_LAUNCHED();
    if ((rc = exclusive_scan_u32(B.flag, B.segid, total + 1, nullptr, B.scan_ws, st))) return rc;
    D3D_CUDA_TRY(cudaMemsetAsync(B.app, 0, (size_t)(total + 1) * 4, st));
    vox_seg_kernel<<<gb, 256, 0, st>>>(B.keys, B.sidx, B.flag, B.segid, total, B.seg_start, B.seg_first, B.app); D3D_LAUNCHED();
    if ((rc = exclusive_scan_u32(B.app, B.app, total + 1, nullptr, B.scan_ws, st))) return rc;
    // the number of segments is data dependent: it lives on the device as segid[total]; every
    // per-segment kernel is launched over total+1 slots and exits on s >= segid[total]
    return D3D_OK;
}

static int check_common(const float *points, int64_t total, int32_t nfeat, const int64_t *offs, int64_t nframes, const d3d_voxel_params *P, int64_t *counts)
{
    if (total < 0 || nframes < 0 || nfeat < 3 || !P || (nframes > 0 && (!offs || !counts))) return D3D_ERR_INVALID_ARGUMENT;
    if (total > 0 && !points) return D3D_ERR_INVALID_ARGUMENT;
    if (total >= (1ll << 31)) return D3D_ERR_INVALID_ARGUMENT;
    if (P->max_points_filter == D3D_PF_FARTHEST_SAMPLING) return D3D_ERR_UNSUPPORTED;   // reference throws too (voxelize.cpp:468-471)
    if (P->max_points_filter < 0 || P->max_points_filter > 2 || P->max_voxels_filter < 0 || P->max_voxels_filter > 2) return D3D_ERR_INVALID_ARGUMENT;
    if (P->reduction < 0 || P->reduction > 3) return D3D_ERR_INVALID_ARGUMENT;
    return D3D_OK;
}

}  // namespace d3d

using namespace d3d;

extern "C" size_t d3d_voxelize_workspace_bytes(int64_t total_points, int64_t nframes, int64_t max_frame_points)
{
    const int64_t t = total_points > 0 ? total_points : 1;
    const size_t a = vox_ws_bytes(t, nframes), b = vox_cluster_ws_bytes(t, nframes, max_frame_points), c = vox_tiles_ws_bytes(t, nframes, max_frame_points);
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

// AUTO: tile pipeline, else cluster path, whenever they support the configuration; an explicit request for an unsupported one is an error
static int pick_algo(const d3d_voxel_params *P, const VoxCfg &cfg, int64_t total, int64_t nframes, bool *cluster, bool *tiles)
{
    if (P->algo < D3D_VOXEL_AUTO || P->algo > D3D_VOXEL_AUTO_NO_TILES) return D3D_ERR_INVALID_ARGUMENT;
    const int64_t mfp = P->max_frame_points > 0 ? P->max_frame_points : total;
    const bool ok = vox_cluster_supported(cfg, total, nframes, mfp);
    const bool okt = vox_tiles_supported(cfg, total, nframes, mfp);
    if (P->algo == D3D_VOXEL_CLUSTER && !ok) return D3D_ERR_UNSUPPORTED;
    if (P->algo == D3D_VOXEL_TILES && !okt) return D3D_ERR_UNSUPPORTED;
    *tiles = okt && (P->algo == D3D_VOXEL_AUTO || P->algo == D3D_VOXEL_TILES);
    *cluster = !*tiles && ok && P->algo != D3D_VOXEL_SORT && P->algo != D3D_VOXEL_TILES;
    return D3D_OK;
}

extern "C" int d3d_voxelize_sparse_f32(const float *points, int64_t total, int32_t nfeat, const int64_t *offs, int64_t nframes, const d3d_voxel_params *P,
                                       float *out_points, int64_t *out_mask, int64_t *out_mapping, int32_t *out_npoints, int64_t *out_coords, int64_t *counts,
                                       void *ws, size_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_common(points, total, nfeat, offs, nframes, P, counts);
    if (rc) return rc;
    if (nframes == 0) return D3D_OK;
    if (total > 0 && (!out_points || !out_mask || !out_mapping || !out_npoints || !out_coords)) return D3D_ERR_INVALID_ARGUMENT;
    VoxCfg cfg;
    if ((rc = build_cfg(P, 0, nframes, &cfg))) return rc;
    bool cluster = false, tiles = false;
    if ((rc = pick_algo(P, cfg, total, nframes, &cluster, &tiles))) return rc;
    if (tiles) {
        if (!ws || ws_bytes < vox_tiles_ws_bytes(total > 0 ? total : 1, nframes, P->max_frame_points)) return D3D_ERR_WORKSPACE;
        return vox_tiles_sparse(points, total, nfeat, offs, nframes, P->max_frame_points, cfg, out_points, out_mask, out_mapping, out_npoints, out_coords,
                                counts, ws, ws_bytes, st);
    }
    if (cluster) {
        if (!ws || ws_bytes < vox_cluster_ws_bytes(total > 0 ? total : 1, nframes, P->max_frame_points)) return D3D_ERR_WORKSPACE;
        return vox_cluster_sparse(points, total, nfeat, offs, nframes, P->max_frame_points, cfg, out_points, out_mask, out_mapping, out_npoints, out_coords,
                                  counts, ws, ws_bytes, st);
    }
    if (!ws || ws_bytes < vox_ws_bytes(total > 0 ? total : 1, nframes)) return D3D_ERR_WORKSPACE;
    if (nframes >= 2048) return D3D_ERR_INVALID_ARGUMENT;   // DESCENDING key packs the frame in 11 bits
    VoxBufs B;
    if ((rc = vox_common(points, total, nfeat, offs, nframes, cfg, B, ws, ws_bytes, st))) return rc;
    const unsigned gb = (unsigned)cdiv(total + 1, 256);
    const uint32_t *nseg = B.segid + total;
    D3D_CUDA_TRY(cudaMemsetAsync(B.passflag, 0, (size_t)(total + 1) * 4, st));
    vox_pass_kernel<<<gb, 256, 0, st>>>(B.keys, B.seg_start, B.seg_first, B.app, nseg, cfg, B.seg_app, B.passflag); D3D_LAUNCHED();
    if ((rc = exclusive_scan_u32(B.passflag, B.prank, total + 1, nullptr, B.scan_ws, st))) return rc;
    if (cfg.vfilter == D3D_VF_DESCENDING) {
        // second stable sort: voxels by (frame, count descending, appearance ascending)
        uint64_t *dk = B.dkeys; uint32_t *dv = B.keepflag;   // keepflag is free until the point pass
        D3D_CUDA_TRY(cudaMemsetAsync(dk, 0xff, (size_t)(total + 1) * 8, st));   // slots past nseg sort last
        vox_desc_key_kernel<<<gb, 256, 0, st>>>(B.keys, B.seg_start, B.seg_app, B.app, offs, nseg, cfg, dk, dv); D3D_LAUNCHED();
        if ((rc = radix_sort_pairs_u64(dk, dv, total + 1, 64, B.sort_ws, B.sort_bytes, st))) return rc;
        vox_desc_rank_kernel<<<gb, 256, 0, st>>>(dk, dv, B.app, offs, nseg, B.desc_rank); D3D_LAUNCHED();
    }
    vox_voxel_rows_kernel<<<1, 1024, 0, st>>>(B.app, B.prank, offs, nframes, cfg, B.voff); D3D_LAUNCHED();
    vox_voxel_out_kernel<<<gb, 256, 0, st>>>(B.keys, B.seg_start, B.seg_app, B.app, B.prank, B.passflag, B.desc_rank, offs, B.voff, nseg, cfg, B.newid, out_npoints, out_coords);
    D3D_LAUNCHED();
    vox_point_flag_kernel<<<gb, 256, 0, st>>>(B.keys, B.sidx, B.segid, B.seg_start, B.newid, total, cfg, B.keepflag, B.pmap); D3D_LAUNCHED();
    D3D_CUDA_TRY(cudaMemsetAsync(B.keepflag + total, 0, 4, st));
    if ((rc = exclusive_scan_u32(B.keepflag, B.ppos, total + 1, nullptr, B.scan_ws, st))) return rc;
    vox_point_out_kernel<<<gb, 256, 0, st>>>(points, nfeat, B.keepflag, B.ppos, B.pmap, offs, nframes, total, out_points, out_mask, out_mapping); D3D_LAUNCHED();
    vox_rows_sparse_kernel<<<(unsigned)cdiv(nframes + 1, 64), 64, 0, st>>>(B.ppos, B.voff, offs, nframes, counts); D3D_LAUNCHED();
    return D3D_OK;
}

extern "C" int d3d_voxelize_dense_f32(const float *points, int64_t total, int32_t nfeat, const int64_t *offs, int64_t nframes, const d3d_voxel_params *P, float *voxels,
                                      int64_t *coords, uint8_t *pmask, int32_t *npoints, float *aggregates, int64_t *counts, void *ws, size_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_common(points, total, nfeat, offs, nframes, P, counts);
    if (rc) return rc;
    if (nframes == 0) return D3D_OK;
    if (P->max_points < 0 || P->max_voxels < 0) return D3D_ERR_INVALID_ARGUMENT;
    if (!coords || !npoints || (P->max_points > 0 && (!voxels || !pmask))) return D3D_ERR_INVALID_ARGUMENT;   // max_points 0: empty slot arrays
    if (P->reduction != D3D_RED_NONE && !aggregates) return D3D_ERR_INVALID_ARGUMENT;
    VoxCfg cfg;
    if ((rc = build_cfg(P, 1, nframes, &cfg))) return rc;
    bool cluster = false, tiles = false;
    if ((rc = pick_algo(P, cfg, total, nframes, &cluster, &tiles))) return rc;
    if (!ws || ws_bytes < (cluster ? vox_cluster_ws_bytes(total > 0 ? total : 1, nframes, P->max_frame_points) : vox_ws_bytes(total > 0 ? total : 1, nframes)))
        return D3D_ERR_WORKSPACE;
    const size_t nslots = (size_t)nframes * (size_t)cfg.max_voxels * (size_t)cfg.max_points;
    D3D_CUDA_TRY(cudaMemsetAsync(voxels, 0, nslots * nfeat * sizeof(float), st));
    D3D_CUDA_TRY(cudaMemsetAsync(pmask, 0, nslots, st));
    D3D_CUDA_TRY(cudaMemsetAsync(npoints, 0, (size_t)nframes * cfg.max_voxels * sizeof(int32_t), st));
    if (cluster)
        return vox_cluster_dense(points, total, nfeat, offs, nframes, P->max_frame_points, cfg, voxels, coords, pmask, npoints, counts, ws, ws_bytes, st);
    VoxBufs B;
    if ((rc = vox_common(points, total, nfeat, offs, nframes, cfg, B, ws, ws_bytes, st))) return rc;
    const unsigned gb = (unsigned)cdiv(total + 1, 256);
    const uint32_t *nseg = B.segid + total;
    D3D_CUDA_TRY(cudaMemsetAsync(B.passflag, 0, (size_t)(total + 1) * 4, st));
    vox_pass_kernel<<<gb, 256, 0, st>>>(B.keys, B.seg_start, B.seg_first, B.app, nseg, cfg, B.seg_app, B.passflag); D3D_LAUNCHED();
    vox_voxel_out_kernel<<<gb, 256, 0, st>>>(B.keys, B.seg_start, B.seg_app, B.app, B.passflag /*unused*/, B.passflag, B.desc_rank, offs, nullptr, nseg, cfg, B.newid, npoints, coords);
    D3D_LAUNCHED();
    vox_dense_point_kernel<<<gb, 256, 0, st>>>(points, nfeat, B.keys, B.sidx, B.segid, B.seg_start, B.newid, total, cfg, voxels, pmask); D3D_LAUNCHED();
    if (cfg.reduction != D3D_RED_NONE) {
        vox_dense_aggr_kernel<<<(unsigned)cdiv((total + 1) * nfeat, 128), 128, 0, st>>>(points, nfeat, B.keys, B.sidx, B.seg_start, B.newid, nseg, cfg, aggregates);
        D3D_LAUNCHED();
    }
    vox_counts_dense_kernel<<<(unsigned)cdiv(nframes, 64), 64, 0, st>>>(B.app, offs, nframes, cfg, counts); D3D_LAUNCHED();
    return D3D_OK;
}
