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

// ---- cluster-per-frame back end (voxel_cluster.cu)
// true when this configuration / problem size is served by the cluster path
bool vox_cluster_supported(const VoxCfg &cfg, int64_t total, int64_t nframes, int64_t max_frame_points);
size_t vox_cluster_ws_bytes(int64_t total, int64_t nframes, int64_t max_frame_points);
int vox_cluster_sparse(const float *points, int64_t total, int nfeat, const int64_t *offs, int64_t nframes, int64_t max_frame_points, const VoxCfg &cfg,
                       float *out_points, int64_t *out_mask, int64_t *out_mapping, int32_t *out_npoints, int64_t *out_coords, int64_t *counts,
                       void *ws, size_t ws_bytes, cudaStream_t st);
int vox_cluster_dense(const float *points, int64_t total, int nfeat, const int64_t *offs, int64_t nframes, int64_t max_frame_points, const VoxCfg &cfg,
                      float *voxels, int64_t *coords, uint8_t *pmask, int32_t *npoints, int64_t *counts, void *ws, size_t ws_bytes, cudaStream_t st);

}  // namespace d3d
