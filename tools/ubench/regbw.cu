// Micro-benchmark (exploration): how many distinct register operands can an FFMA2 / FFMA read per cycle on sm_100a?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 regbw.cu -o regbw
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define N 12
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, const float* in) {
    float2 acc[N], v[N], w[N];
    float sc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        acc[i] = make_float2(in[i] + threadIdx.x, in[i + 1]); v[i] = make_float2(in[i + 2] * 0.5f, in[i + 3] * 0.25f);
        w[i] = make_float2(in[i + 4] * 0.125f, in[i + 5]); sc[i] = in[i + 6] * 1e-3f;
    }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE == 0) acc[i] = __ffma2_rn(make_float2(sc[i], sc[i]), v[0], acc[i]);            // scalar_i, pair const, acc_i : 3 new regs
            if (MODE == 1) acc[i] = __ffma2_rn(make_float2(sc[i], sc[i]), v[i], acc[i]);            // scalar_i, pair_i, acc_i   : 5 new regs
            if (MODE == 2) acc[i] = __ffma2_rn(w[i], v[i], acc[i]);                                  // pair_i, pair_i, acc_i     : 6 new regs
            if (MODE == 3) acc[i] = __ffma2_rn(make_float2(sc[0], sc[0]), v[i], acc[i]);            // scalar const, pair_i, acc_i: 4 new regs
            if (MODE == 4) acc[i] = __ffma2_rn(make_float2(sc[i & ~1], sc[i & ~1]), v[i], acc[i]);  // scalar shared by 2 consecutive: 4.5 avg
            if (MODE == 5) { acc[i].x = fmaf(sc[i], v[i].x, acc[i].x); acc[i].y = fmaf(sc[i], v[i].y, acc[i].y); }   // 2 FFMA, 3 regs each (scalar reused)
            if (MODE == 6) { acc[i].x = fmaf(w[i].x, v[i].x, acc[i].x); acc[i].y = fmaf(w[i].y, v[i].y, acc[i].y); } // 2 FFMA, 3 new regs each
            if (MODE == 7) acc[i] = __ffma2_rn(make_float2(sc[i], sc[i]), v[i & 1], acc[i]);        // scalar_i, pair alternating between 2, acc_i
            if (MODE == 8) acc[i] = __ffma2_rn(v[0], w[0], acc[i]);                                  // only acc varies: 2 new regs
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* out, const float* in) {
    k<MODE><<<296, 256>>>(out, in);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<296, 256>>>(out, in); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double packed = 296.0 * 8 * (double)ITERS * N;   // warp-level packed-FMA (or FFMA pair) instructions
    // per SMSP: 148*4 schedulers
    printf("%-46s %.3f ms  %.2f cycles per packed FMA per SMSP @1.965GHz  (%.1f TFLOP/s)\n", name, ms, ms * 1e-3 * 1.965e9 * 592 / packed, packed * 32 * 4 / (ms * 1e-3) / 1e12);
}
int main() {
    float *out, *in; cudaMalloc(&out, 296 * 256 * 4); cudaMalloc(&in, 64 * 4);
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f + 0.01f * i; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<8>("FFMA2 const,const,acc_i (2 regs)", out, in);
    run<0>("FFMA2 scalar_i, pair const, acc_i (3 regs)", out, in);
    run<3>("FFMA2 scalar const, pair_i, acc_i (4 regs)", out, in);
    run<4>("FFMA2 scalar shared by 2, pair_i, acc_i", out, in);
    run<7>("FFMA2 scalar_i, pair alt 2, acc_i", out, in);
    run<1>("FFMA2 scalar_i, pair_i, acc_i (5 regs)", out, in);
    run<2>("FFMA2 pair_i, pair_i, acc_i (6 regs)", out, in);
    run<5>("2xFFMA scalar_i, v_i, acc_i", out, in);
    run<6>("2xFFMA w_i, v_i, acc_i", out, in);
    return 0;
}
