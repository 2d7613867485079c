// tuning probe (not part of the product): issue cost of warp-level primitives on sm_100a, cycles per warp instruction per SM
// (all warps of full-occupancy CTAs issue the same primitive back to back).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o warp_prims_probe warp_prims_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
template <int OP> __global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t *gmem, long long *cyc)
{
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
    const unsigned lane = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; i++) {
        x = x * 1664525u + 1013904223u;
        if (OP == 0) acc += __match_any_sync(0xffffffffu, x >> 27);                 // 32 values
        if (OP == 1) acc += __ballot_sync(0xffffffffu, x >> 31);
        if (OP == 2) acc += __shfl_sync(0xffffffffu, x, (x >> 27));
        if (OP == 3) acc += atomicAdd(&sm[(x >> 20)], 1u);                          // random word of 4096, with return
        if (OP == 4) atomicAdd(&sm[(x >> 20)], 1u);                                 // no return
        if (OP == 5) acc += atomicCAS(&sm[(x >> 20)], 0xffffffffu, x);
        if (OP == 6) { sm[(x >> 20)] = x; acc += sm[((x >> 8) & 4095)]; }            // plain store + load, random
        if (OP == 7) acc += __match_any_sync(0xffffffffu, x);                       // all distinct
        if (OP == 8) acc += __popc(__activemask() & (x | 1));                       // baseline ALU
        if (OP == 9) acc += atomicAdd(&gmem[(x >> 12)], 1u);                        // global, 1M words, with return
        if (OP == 10) atomicAdd(&gmem[(x >> 12)], 1u);                              // global RED
        if (OP == 11) acc += __reduce_add_sync(0xffffffffu, x);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + lane;
}
template <int OP> void run(const char *name, uint32_t *out, uint32_t *gmem, long long *cyc, int sms)
{
    k<OP><<<sms * 2, 1024>>>(out, gmem, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<OP><<<sms * 2, 1024>>>(out, gmem, cyc); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[2]; cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    // 64 warps per SM each issue ITERS primitives: cycles per warp-instruction per SM
    printf("%-44s %8.2f cyc per warp instr per SM   (%.3f ms, %s)\n", name, (double)h[0] / (ITERS * 64.0), ms, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out, *gmem; long long *cyc;
    cudaMalloc(&out, (size_t)sms * 2 * 1024 * 4); cudaMalloc(&gmem, 4 << 20); cudaMemset(gmem, 0, 4 << 20); cudaMalloc(&cyc, sms * 2 * 8);
    run<8>("baseline (lcg + popc)", out, gmem, cyc, sms);
    run<0>("match.any, 32 values", out, gmem, cyc, sms);
    run<7>("match.any, all distinct", out, gmem, cyc, sms);
    run<1>("vote.ballot", out, gmem, cyc, sms);
    run<2>("shfl.idx", out, gmem, cyc, sms);
    run<11>("redux.add", out, gmem, cyc, sms);
    run<3>("shared atomicAdd with result, random", out, gmem, cyc, sms);
    run<4>("shared atomicAdd no result, random", out, gmem, cyc, sms);
    run<5>("shared atomicCAS, random", out, gmem, cyc, sms);
    run<6>("shared plain store + load, random", out, gmem, cyc, sms);
    run<9>("global atomicAdd with result, 1M words", out, gmem, cyc, sms);
    run<10>("global atomicAdd no result, 1M words", out, gmem, cyc, sms);
    return 0;
}
