// Micro-benchmark: scalar FFMA vs packed FFMA2/FMUL2/FADD2 issue throughput and latency on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/ubench/ffma2 tools/ubench/ffma2.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE, int CH>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b)
{
    float2 acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) {  // 2 scalar FFMA
                acc[i].x = __fmaf_rn(acc[i].x, A.x, B.x);
                acc[i].y = __fmaf_rn(acc[i].y, A.y, B.y);
            } else if (MODE == 1) {  // 1 FFMA2
                acc[i] = __ffma2_rn(acc[i], A, B);
            } else if (MODE == 2) {  // FMUL2
                acc[i] = __fmul2_rn(acc[i], A);
            } else if (MODE == 3) {  // FADD2
                acc[i] = __fadd2_rn(acc[i], B);
            } else if (MODE == 4) {  // FFMA2 + independent integer ALU work (co-issue test)
                acc[i] = __ffma2_rn(acc[i], A, B);
                acc[i].x = __int_as_float(__float_as_int(acc[i].x) ^ 1);
            } else if (MODE == 5) {  // scalar FFMA x2 + same ALU op
                acc[i].x = __fmaf_rn(acc[i].x, A.x, B.x);
                acc[i].y = __fmaf_rn(acc[i].y, A.y, B.y);
                acc[i].x = __int_as_float(__float_as_int(acc[i].x) ^ 1);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i].x + acc[i].y;
    if (s == 123.456f) out[0] = s;
}

template <int MODE, int CH>
void run(const char *name, int sms, float clk_ghz)
{
    float *d;
    cudaMalloc(&d, 4);
    const int iters = 20000, grid = sms * 8;
    k<MODE, CH><<<grid, 256>>>(d, 100, 1.0001f, 1e-3f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, CH><<<grid, 256>>>(d, iters, 1.0001f, 1e-3f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_fma = (double)grid * 256 * iters * CH * 2;  // scalar-equivalent fp ops (x and y)
    double per_clk_sm = lane_fma / (ms * 1e-3) / (clk_ghz * 1e9) / sms;
    printf("%-28s CH=%d  %8.3f ms  %7.1f fp32 lane-ops/clk/SM (128 = scalar peak)\n", name, CH, ms, per_clk_sm);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float ghz = khz / 1e6f;
    printf("%s SMs %d clock %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    int sms = p.multiProcessorCount;
    run<0, 8>("FFMA scalar", sms, ghz);
    run<1, 8>("FFMA2 packed", sms, ghz);
    run<2, 8>("FMUL2 packed", sms, ghz);
    run<3, 8>("FADD2 packed", sms, ghz);
    run<4, 8>("FFMA2 + LOP", sms, ghz);
    run<5, 8>("FFMA,FFMA + LOP", sms, ghz);
    run<0, 1>("FFMA scalar (latency)", sms, ghz);
    run<1, 1>("FFMA2 packed (latency)", sms, ghz);
    run<0, 2>("FFMA scalar", sms, ghz);
    run<1, 2>("FFMA2 packed", sms, ghz);
    return 0;
}
