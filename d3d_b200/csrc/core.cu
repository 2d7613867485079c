// core.cu -- library-wide state of libd3d_b200.so: error reporting, launch counter, ALU-peak probe.
#include "common.cuh"
#include <string.h>

namespace d3d {
std::atomic<int64_t> g_launches{0};
static thread_local char g_cuda_err[256] = "";
void set_cuda_error(cudaError_t e)
{
    strncpy(g_cuda_err, cudaGetErrorString(e), sizeof(g_cuda_err) - 1);
    g_cuda_err[sizeof(g_cuda_err) - 1] = 0;
}

// FMA-chain microbenchmark: 8 independent chains per thread, `iters` rounds -> measured CUDA-core peak
template <typename T>
__global__ void __launch_bounds__(256) fma_probe_kernel(int64_t iters, T seed, float *sink)
{
    T a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = T(0.999), c = T(0.001);
    for (int64_t i = 0; i < iters; i++) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == T(-1)) sink[0] = (float)s;  // never true; keeps the chains alive
}
}  // namespace d3d

using namespace d3d;

extern "C" int d3d_abi_version(void) { return D3D_B200_ABI_VERSION; }
extern "C" int64_t d3d_launch_count(void) { return g_launches.load(); }
extern "C" const char *d3d_last_cuda_error(void) { return g_cuda_err; }
extern "C" const char *d3d_error_string(int s)
{
    switch (s) {
    case D3D_OK: return "ok";
    case D3D_ERR_INVALID_ARGUMENT: return "invalid argument";
    case D3D_ERR_CUDA: return "CUDA error";
    case D3D_ERR_WORKSPACE: return "workspace too small";
    case D3D_ERR_UNSUPPORTED: return "unsupported option";
    case D3D_ERR_RANGE: return "voxel grid exceeds the 63-bit key range";
    default: return "unknown status";
    }
}
extern "C" int d3d_fma_peak_probe(int dtype, int64_t iters, float *sink, double *flops_host, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    D3D_CUDA_TRY(cudaGetDevice(&dev));
    D3D_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256;   // 2048 threads per SM: every scheduler saturated
    if (dtype == D3D_F64) fma_probe_kernel<double><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else fma_probe_kernel<float><<<blocks, threads, 0, st>>>(iters, 1.0f, sink);
    D3D_LAUNCHED();
    if (flops_host) *flops_host = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    return D3D_OK;
}
