// Micro-benchmark (exploration, not product): issue/pipe behaviour of FFMA, FFMA2 (packed FP32), ALU ops and
// MUFU on sm_100a, to size the filter loop of the scan kernels.  Prints cycles per loop iteration per warp
// scheduler with 1..8 warps per SMSP.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes.cu -o pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void k(float* out, unsigned long long* cyc, float seed) {
    float2 a[8];
    float f[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed + i, seed - i); f[i] = seed * i; u[i] = threadIdx.x * (i + 1); }
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { f[i] = fmaf(f[i], m.x, c.x); a[i].x = fmaf(a[i].x, m.y, c.y); }                       // 16 FFMA
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); }                                                          // 8 FFMA2 (=16 FMA/lane)
            if (MODE == 2) { a[i] = __ffma2_rn(a[i], m, c); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e37u; }                   // 8 FFMA2 + ~16 ALU
            if (MODE == 3) { f[i] = fmaf(f[i], m.x, c.x); a[i].x = fmaf(a[i].x, m.y, c.y); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e37u; }  // 16 FFMA + ~16 ALU
            if (MODE == 4) { a[i] = __ffma2_rn(a[i], m, c); u[i] = u[i] + 0x9e37u; }                                   // 8 FFMA2 + 8 IADD
            if (MODE == 5) { a[i] = __ffma2_rn(a[i], m, c); if (i < 2) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f[i])); f[i] = r; } }  // 8 FFMA2 + 2 MUFU
            if (MODE == 6) { u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e37u; }                                                  // ALU only
            if (MODE == 7) { a[i] = __ffma2_rn(a[i], m, c); f[i] = fminf(f[i], a[(i + 1) & 7].x); }                    // 8 FFMA2 + 8 FMNMX
            if (MODE == 8) { a[i] = __ffma2_rn(a[i], m, c); a[(i + 3) & 7].y = fmaf(a[(i + 3) & 7].y, m.x, c.x); }     // 8 FFMA2 + 8 FFMA
        }
    }
    unsigned long long t1 = clock64();
    float s = 0; unsigned v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += a[i].x + a[i].y + f[i]; v ^= u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + v;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, unsigned long long* cyc) {
    printf("%-28s", name);
    for (int warps_per_smsp : {1, 2, 4, 8}) {
        int threads = 32 * 4 * warps_per_smsp;  // one CTA per SM, warps spread over the 4 SMSPs
        k<MODE><<<148, threads>>>(out, cyc, 1.0f);
        cudaDeviceSynchronize();
        unsigned long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  w%d: %6.2f cyc/iter/SMSP", warps_per_smsp, (double)h / ITERS);
    }
    printf("\n");
}

int main() {
    float* out; unsigned long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    printf("per iteration (8x unrolled body); cycles are per SMSP for ALL its warps' iterations = wall cycles / ITERS\n");
    run<0>("16 FFMA", out, cyc);
    run<1>("8 FFMA2", out, cyc);
    run<2>("8 FFMA2 + 8x(LOP3,SHF,IADD)", out, cyc);
    run<3>("16 FFMA + 8x(LOP3,SHF,IADD)", out, cyc);
    run<4>("8 FFMA2 + 8 IADD", out, cyc);
    run<5>("8 FFMA2 + 2 MUFU.RCP", out, cyc);
    run<6>("8x(LOP3,SHF,IADD) only", out, cyc);
    run<7>("8 FFMA2 + 8 FMNMX", out, cyc);
    run<8>("8 FFMA2 + 8 FFMA", out, cyc);
    return 0;
}
