// What SM clock does a dense FMA loop really run at?  cycles (clock64) / event time.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(float2* out, unsigned long long* cyc, int iters, int packed) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(1.0f + i, 1.0f - i + threadIdx.x * 1e-3f);
    const float2 m = make_float2(0.9999f, 1.0001f), c = make_float2(1e-4f, -1e-4f);
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (packed) acc[i] = __ffma2_rn(acc[i], m, c);
            else { acc[i].x = fmaf(acc[i].x, m.x, c.x); acc[i].y = fmaf(acc[i].y, m.y, c.y); }
        }
    }
    unsigned long long t1 = clock64();
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    const int sms = 148;
    float2* out; unsigned long long* cyc; cudaMalloc(&out, sizeof(float2) * sms * 8 * 256); cudaMalloc(&cyc, 8 * sms * 8);
    for (int ctas_per_sm : {1, 2, 4, 8}) for (int packed : {1, 0}) {
        const int grid = sms * ctas_per_sm, iters = 16384;
        k<<<grid, 256>>>(out, cyc, iters / 8, packed);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<grid, 256>>>(out, cyc, iters, packed); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h[8 * 148]; cudaMemcpy(h, cyc, 8 * grid, cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
        const double flop = (double)grid * 256 * (double)iters * 16 * 4;
        printf("ctas/SM %d %s: %.3f ms, %.1f TFLOP/s, mean CTA cycles %.0f -> SM clock %.0f MHz, cycles per warp-FFMA2-equivalent per SMSP %.3f\n", ctas_per_sm,
               packed ? "FFMA2" : "FFMA ", ms, flop / (ms * 1e-3) / 1e12, mean, mean / (ms * 1e3), mean / ((double)iters * 16 * ctas_per_sm * 2));
    }
    return 0;
}
