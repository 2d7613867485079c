// tuning probe (not product): cost of distributed-shared-memory stores by access pattern, cluster of 8 x 1024 threads
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dsmem_probe tools/dsmem_probe.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;
constexpr int WORDS = 16384;   // 64 KB target buffer per CTA
constexpr int ITER = 64;

__device__ __forceinline__ uint32_t mix(uint32_t x) { x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA6Bu; x ^= x >> 13; return x; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(long long *out, uint32_t *sink)
{
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) uint32_t buf[];
    const unsigned cs = cluster.num_blocks(), cr = cluster.block_rank(), tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < WORDS; i += 1024) buf[i] = i;
    cluster.sync();
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITER; it++) {
        const uint32_t h = mix(tid * 131u + it * 7919u + cr * 977u);
        if (MODE == 0) {          // 4 B, random CTA, random word
            cluster.map_shared_rank(buf, h % cs)[(h >> 8) % WORDS] = h;
        } else if (MODE == 1) {   // 8 B, random CTA, random word pair
            reinterpret_cast<uint2 *>(cluster.map_shared_rank(buf, h % cs))[(h >> 8) % (WORDS / 2)] = make_uint2(h, h);
        } else if (MODE == 2) {   // 4 B, one CTA per warp instruction, 128 contiguous bytes
            const uint32_t hw = mix(w * 131u + it * 7919u + cr * 977u);
            cluster.map_shared_rank(buf, hw % cs)[((hw >> 8) % (WORDS / 32)) * 32 + lane] = h;
        } else if (MODE == 3) {   // 8 B, runs of 4 lanes to one CTA, 32 contiguous bytes per run
            const uint32_t hr = mix((tid >> 2) * 131u + it * 7919u + cr * 977u);
            reinterpret_cast<uint2 *>(cluster.map_shared_rank(buf, hr % cs))[((hr >> 8) % (WORDS / 8)) * 4 + (lane & 3)] = make_uint2(h, h);
        } else if (MODE == 4) {   // 16 B, one CTA per warp instruction, 512 contiguous bytes
            const uint32_t hw = mix(w * 131u + it * 7919u + cr * 977u);
            reinterpret_cast<uint4 *>(cluster.map_shared_rank(buf, hw % cs))[((hw >> 8) % (WORDS / 128)) * 32 + lane] = make_uint4(h, h, h, h);
        } else if (MODE == 5) {   // local 4 B random store
            buf[(h >> 8) % WORDS] = h;
        } else if (MODE == 6) {   // 4 B random remote load
            acc += cluster.map_shared_rank(buf, h % cs)[(h >> 8) % WORDS];
        } else if (MODE == 7) {   // remote atomicAdd, random
            atomicAdd(cluster.map_shared_rank(buf, h % cs) + (h >> 8) % WORDS, 1u);
        } else if (MODE == 8) {   // local atomicAdd, random
            atomicAdd(buf + (h >> 8) % WORDS, 1u);
        } else if (MODE == 9) {   // 8 B, runs of 12 lanes (as a dest-sorted batch would give): lanes 0-11, 12-23, 24-31
            const uint32_t run = lane / 12;
            const uint32_t hr = mix((w * 3 + run) * 131u + it * 7919u + cr * 977u);
            reinterpret_cast<uint2 *>(cluster.map_shared_rank(buf, hr % cs))[((hr >> 8) % (WORDS / 32)) * 16 + (lane % 12)] = make_uint2(h, h);
        } else if (MODE == 10) {  // local 4 B random load
            acc += buf[(h >> 8) % WORDS];
        }
    }
    cluster.sync();
    const long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) sink[0] = acc;
}

template <int MODE>
static void run(const char *name, long long *d_out, uint32_t *sink)
{
    auto k = probe<MODE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4);
    cudaLaunchConfig_t lc = {};
    lc.blockDim = dim3(1024); lc.gridDim = dim3(8 * 8); lc.dynamicSmemBytes = WORDS * 4;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    long long h[64];
    for (int rep = 0; rep < 2; rep++) { cudaLaunchKernelEx(&lc, k, d_out, sink); cudaDeviceSynchronize(); }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 64; i++) s += (double)h[i];
    s /= 64;
    printf("%-52s %8.0f cyc/CTA  %6.2f cyc per warp instruction  %5.2f cyc/lane  (%s)\n", name, s, s / (32.0 * ITER), s / (1024.0 * ITER), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    long long *d_out; uint32_t *sink;
    cudaMalloc(&d_out, 64 * 8); cudaMalloc(&sink, 4);
    run<0>("remote store 4B random cta/addr", d_out, sink);
    run<1>("remote store 8B random cta/addr", d_out, sink);
    run<2>("remote store 4B warp->one cta, 128B contiguous", d_out, sink);
    run<3>("remote store 8B runs of 4 lanes (32B)", d_out, sink);
    run<9>("remote store 8B runs of 12 lanes (96B)", d_out, sink);
    run<4>("remote store 16B warp->one cta, 512B contiguous", d_out, sink);
    run<5>("local store 4B random", d_out, sink);
    run<10>("local load 4B random", d_out, sink);
    run<6>("remote load 4B random", d_out, sink);
    run<7>("remote atomicAdd random", d_out, sink);
    run<8>("local atomicAdd random", d_out, sink);
    return 0;
}
