// voxel.cuh -- configuration shared by the two voxelization back ends: the cluster-per-frame hash path
// (voxel_cluster.cu, the fast path) and the sort/segmented-scan path (voxel.cu, the general fallback).
#pragma once
#include "common.cuh"

namespace d3d {

constexpr uint64_t VOX_INVALID = ~0ull;

struct VoxCfg {
    int dense;
    float size[3];      // voxel size (sparse: _size tensor; dense: (hi-lo)/shape in float)
    float lo[3];        // dense lower bound
    long long vlo[3];   // sparse: first kept coordinate; dense: 0
    long long ext[3];   // kept extent per dim
    int offset[3];      // sparse: coords_out = coord - offset
    unsigned long long G;  // cells per frame
    int min_points, max_points, max_voxels, pfilter, vfilter, reduction;
};

// grid cell of one point with the reference's exact fp32 arithmetic; false when the point falls outside
// the kept extent (or is NaN).  lin = ((cx * ext_y) + cy) * ext_z + cz with c relative to vlo.
__device__ __forceinline__ bool vox_cell(const VoxCfg &cfg, float px, float py, float pz, unsigned long long *lin_out)
{
    const float p[3] = {px, py, pz};
    bool ok = true;
    unsigned long long lin = 0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        long long c;
        if (cfg.dense) {
            float v = __fdiv_rn(__fsub_rn(p[d], cfg.lo[d]), cfg.size[d]);   // voxelize.cpp:100, truncation toward zero
            ok = ok && !isnan(v);
            c = (long long)(int)v;
        } else {
            float v = floorf(__fdiv_rn(p[d], cfg.size[d]));                 // voxelize.cpp:309
            ok = ok && !isnan(v);
            c = (long long)(int)v - cfg.vlo[d];
        }
        ok = ok && c >= 0 && c < cfg.ext[d];
        lin = lin * (unsigned long long)cfg.ext[d] + (unsigned long long)(ok ? c : 0);
    }
    *lin_out = lin;
    return ok;
}

// ---- 32-bit cell keys shared by the cluster path (voxel_cluster.cu) and the tile pipeline (voxel_tiles.cu)
constexpr uint32_t VC_NOKEY = 0xffffffffu;   // cell keys use at most 31 bits
// per-launch constants in the 32-bit form the kernel computes with
struct VcDev {
    float size[3], lo[3];
    int vlo[3];
    uint32_t ext[3];
    uint32_t sh_x, sh_y;     // cell key = cx << sh_x | cy << sh_y | cz (bit fields: decoding is two shifts and two masks)
    long long cadd[3];       // coords_out = c + vlo - offset
};

template <bool DENSE>
__device__ __forceinline__ bool vc_cell(const VcDev &c, const float4 &p, uint32_t *key)
{
    float v0, v1, v2;
    int i0, i1, i2;
    if (DENSE) {
        v0 = __fdiv_rn(__fsub_rn(p.x, c.lo[0]), c.size[0]);
        v1 = __fdiv_rn(__fsub_rn(p.y, c.lo[1]), c.size[1]);
        v2 = __fdiv_rn(__fsub_rn(p.z, c.lo[2]), c.size[2]);
        i0 = (int)v0; i1 = (int)v1; i2 = (int)v2;                       // truncation toward zero, voxelize.cpp:100
    } else {
        v0 = __fdiv_rn(p.x, c.size[0]);
        v1 = __fdiv_rn(p.y, c.size[1]);
        v2 = __fdiv_rn(p.z, c.size[2]);
        // (int)floorf(v) in one conversion (F2I.FLOOR): same value for every finite v, same saturation beyond the int range; NaN is
        // rejected below.  One trip through the conversion unit per coordinate instead of two (the split kernel is bound by it)
        i0 = __float2int_rd(v0); i1 = __float2int_rd(v1); i2 = __float2int_rd(v2);
    }
    const uint32_t c0 = (uint32_t)(i0 - c.vlo[0]), c1 = (uint32_t)(i1 - c.vlo[1]), c2 = (uint32_t)(i2 - c.vlo[2]);
    *key = (c0 << c.sh_x) | (c1 << c.sh_y) | c2;
    return v0 == v0 && v1 == v1 && v2 == v2 && c0 < c.ext[0] && c1 < c.ext[1] && c2 < c.ext[2];
}

static inline int vox_bits_for(long long ext)   // bits needed for coordinates 0 .. ext-1
{
    int b = 0;
    while ((1ll << b) < ext) b++;
    return b;
}

static inline bool vc_make_dev(const VoxCfg &cfg, VcDev *d)
{
    int bits[3];
    for (int k = 0; k < 3; k++) {
        if (cfg.ext[k] <= 0 || cfg.ext[k] > (1ll << 30)) return false;
        if (cfg.vlo[k] <= -(1ll << 30) || cfg.vlo[k] >= (1ll << 30)) return false;   // 32-bit cell arithmetic in vc_cell
        bits[k] = vox_bits_for(cfg.ext[k]);
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

// ---- cluster-per-frame back end (voxel_cluster.cu)
// true when this configuration / problem size is served by the cluster path
bool vox_cluster_supported(const VoxCfg &cfg, int64_t total, int64_t nframes, int64_t max_frame_points);
size_t vox_cluster_ws_bytes(int64_t total, int64_t nframes, int64_t max_frame_points);
int vox_cluster_sparse(const float *points, int64_t total, int nfeat, const int64_t *offs, int64_t nframes, int64_t max_frame_points, const VoxCfg &cfg,
                       float *out_points, int64_t *out_mask, int64_t *out_mapping, int32_t *out_npoints, int64_t *out_coords, int64_t *counts,
                       void *ws, size_t ws_bytes, cudaStream_t st, const uint32_t *run_if = nullptr);
int vox_cluster_dense(const float *points, int64_t total, int nfeat, const int64_t *offs, int64_t nframes, int64_t max_frame_points, const VoxCfg &cfg,
                      float *voxels, int64_t *coords, uint8_t *pmask, int32_t *npoints, int64_t *counts, void *ws, size_t ws_bytes, cudaStream_t st);

// ---- tile pipeline (voxel_tiles.cu): sparse, no voxel cap, TRIM up to 8 points per voxel
bool vox_tiles_supported(const VoxCfg &cfg, int64_t total, int64_t nframes, int64_t max_frame_points);
size_t vox_tiles_ws_bytes(int64_t total, int64_t nframes, int64_t max_frame_points);
int vox_tiles_sparse(const float *points, int64_t total, int nfeat, const int64_t *offs, int64_t nframes, int64_t max_frame_points, const VoxCfg &cfg,
                     float *out_points, int64_t *out_mask, int64_t *out_mapping, int32_t *out_npoints, int64_t *out_coords, int64_t *counts,
                     void *ws, size_t ws_bytes, cudaStream_t st);

}  // namespace d3d
