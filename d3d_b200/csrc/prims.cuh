// prims.cuh -- device-wide scan / radix sort used by NMS and voxelization (implemented in prims.cu)
#pragma once
#include "common.cuh"

namespace d3d {
size_t scan_workspace_bytes(int64_t n);
// out[i] = sum(in[0..i)); in == out allowed.  total_out (device, optional) receives sum(in[0..n)).
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total_out, void *ws, cudaStream_t st);
size_t radix_sort_workspace_bytes(int64_t n);
// stable ascending sort of (key, val) on the low key_bits bits of key
int radix_sort_pairs_u64(uint64_t *keys, uint32_t *vals, int64_t n, int key_bits, void *ws, size_t ws_bytes, cudaStream_t st);
// the same order for full 64-bit keys whose low half rarely decides (double-precision scores): four passes on the high half + a fix-up of
// the runs of equal high halves; `flag` is one device word of scratch
int radix_sort_pairs_u64_hi32(uint64_t *keys, uint32_t *vals, int64_t n, void *ws, size_t ws_bytes, uint32_t *flag, cudaStream_t st);
}  // namespace d3d
