// write_bw_probe.cu -- what a write-only stream reaches on this GPU next to a read-only stream and a copy (the roofline denominator in
// MEASURED_PEAKS.json is a copy: half reads, half writes).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_bw_probe write_bw_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(256) store_kernel(uint4 *p, size_t n4)
{
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (MODE == 0) p[i] = z;
        if (MODE == 1) __stcs(p + i, z);
        if (MODE == 2) __stwt(p + i, z);
    }
}
// one CTA per contiguous piece (the shape of the crop rows kernel)
__global__ void __launch_bounds__(256) piece_kernel(uint4 *p, size_t piece4)
{
    uint4 *q = p + (size_t)blockIdx.x * piece4;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = threadIdx.x; i < piece4; i += blockDim.x) q[i] = z;
}
__global__ void __launch_bounds__(256) read_kernel(const uint4 *p, size_t n4, uint32_t *out)
{
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { const uint4 v = __ldcs(p + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *out = acc;
}
__global__ void __launch_bounds__(256) copy_kernel(const uint4 *a, uint4 *b, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(b + i, __ldcs(a + i));
}

template <typename F> static float timeit(F f, int reps = 10)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main()
{
    const size_t bytes = (size_t)1 << 30, n4 = bytes / 16;
    uint4 *a, *b; uint32_t *o;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 4);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    float ms;
    ms = timeit([&] { cudaMemsetAsync(a, 0, bytes); });
    printf("cudaMemsetAsync 1 GiB              %8.3f ms  %7.1f GB/s written\n", ms, bytes / ms / 1e6);
    for (int g : {148 * 2, 148 * 8, 148 * 32}) {
        ms = timeit([&] { store_kernel<0><<<g, 256>>>(a, n4); });
        printf("st.global.v4   grid %5d          %8.3f ms  %7.1f GB/s written\n", g, ms, bytes / ms / 1e6);
        ms = timeit([&] { store_kernel<1><<<g, 256>>>(a, n4); });
        printf("st.global.cs   grid %5d          %8.3f ms  %7.1f GB/s written\n", g, ms, bytes / ms / 1e6);
        ms = timeit([&] { store_kernel<2><<<g, 256>>>(a, n4); });
        printf("st.global.wt   grid %5d          %8.3f ms  %7.1f GB/s written\n", g, ms, bytes / ms / 1e6);
    }
    for (size_t piece : {(size_t)16384, (size_t)61440, (size_t)262144}) {
        ms = timeit([&] { piece_kernel<<<(unsigned)(bytes / piece), 256>>>(a, piece / 16); });
        printf("one CTA per %6zu-byte piece       %8.3f ms  %7.1f GB/s written\n", piece, ms, (bytes / piece * piece) / ms / 1e6);
    }
    ms = timeit([&] { read_kernel<<<148 * 16, 256>>>(a, n4, o); });
    printf("read only (ld.global.cs.v4)         %8.3f ms  %7.1f GB/s read\n", ms, bytes / ms / 1e6);
    ms = timeit([&] { copy_kernel<<<148 * 16, 256>>>(a, b, n4); });
    printf("copy kernel                         %8.3f ms  %7.1f GB/s read + written\n", ms, 2.0 * bytes / ms / 1e6);
    ms = timeit([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpyAsync D2D                 %8.3f ms  %7.1f GB/s read + written\n", ms, 2.0 * bytes / ms / 1e6);
    return 0;
}
