// fp32x2_probe.cu -- tuning probe (not part of the product): issue rate of scalar FFMA vs packed FFMA2 /
// FADD2 / FMUL2 on sm_100a, alone and interleaved with ALU-pipe FMNMX, to decide whether the IoU clip
// should be written on float2 lanes.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_probe fp32x2_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
constexpr int CH = 8, IT = 4096;
template <int MODE> __global__ void __launch_bounds__(512) k(float *out, float a, float b)
{
    float2 x[CH]; float m[CH];
    for (int i = 0; i < CH; i++) { x[i] = make_float2(threadIdx.x + i, threadIdx.x - i); m[i] = i; }
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
    for (int it = 0; it < IT; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a2.y, b2.y); }          // 2 FFMA
            if (MODE == 1) x[i] = __ffma2_rn(x[i], a2, b2);                                              // 1 FFMA2
            if (MODE == 2) { x[i] = __ffma2_rn(x[i], a2, b2); m[i] = fminf(fmaxf(m[i], x[i].x), b); }      // FFMA2 + 2 FMNMX
            if (MODE == 3) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a2.y, b2.y); m[i] = fminf(fmaxf(m[i], x[i].x), b); }  // 2 FFMA + 2 FMNMX
            if (MODE == 4) { x[i] = __fadd2_rn(x[i], a2); x[i] = __fmul2_rn(x[i], b2); }                    // FADD2 + FMUL2
            if (MODE == 5) { x[i].x = (x[i].x + a) * b; x[i].y = (x[i].y + a2.y) * b2.y; }                  // 2 FADD + 2 FMUL (no contraction wanted)
        }
    }
    float s = 0;
    for (int i = 0; i < CH; i++) s += x[i].x + x[i].y + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double flop_per_iter_chain, double inst_per_iter_chain)
{
    int dev = 0, sms = 0, khz = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float *out; cudaMalloc(&out, (size_t)sms * 4 * 512 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) k<MODE><<<sms * 4, 512>>>(out, 1.0001f, 0.0001f);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; r++) k<MODE><<<sms * 4, 512>>>(out, 1.0001f, 0.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double threads = (double)sms * 4 * 512, work = threads * IT * CH;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  %6.3f warp-inst/clk/SMSP (at %d MHz nominal)\n", name, ms, work * flop_per_iter_chain / ms * 1e-9,
           work / 32 * inst_per_iter_chain / (ms * 1e-3) / ((double)khz * 1e3) / (sms * 4), khz / 1000);
    cudaFree(out);
}
int main()
{
    run<0>("2xFFMA", 4, 2); run<1>("FFMA2", 4, 1); run<2>("FFMA2+2FMNMX", 4, 3); run<3>("2FFMA+2FMNMX", 4, 4);
    run<4>("FADD2+FMUL2", 4, 2); run<5>("2FADD+2FMUL", 4, 4);
    return 0;
}
