// Micro-benchmark (exploration, not product): the pencil filter's hot loop on sm_100a -- packed (FFMA2) vs scalar
// (FFMA) evaluation of the three weights, rays per thread, triangles per block, resident CTAs.  One 128-triangle tile of
// records sits in shared memory and is scanned ITERS times by every warp; prints SMSP cycles per (ray pair, triangle).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo pencil.cu -o pencil
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kTile = 128;
#define ITERS 256

__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }

template <int RP> struct Rays { float2 x[RP], y[RP]; };

// V0: shipped -- packed, two triangles per AND
template <int RP, int J>
__device__ __forceinline__ uint32_t block_packed(const Rays<RP>& f, const float4* rec) {
    uint32_t acc[2 * RP];
#pragma unroll
    for (int k = 0; k < 2 * RP; ++k) acc[k] = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q0 = rec[j * 4 + 0], q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float2 a = __ffma2_rn(splat2(q0.y), f.y[p], splat2(q0.z));
            float2 b = __ffma2_rn(splat2(q1.y), f.y[p], splat2(q1.z));
            float2 c = __ffma2_rn(splat2(q2.y), f.y[p], splat2(q2.z));
            a = __ffma2_rn(splat2(q0.x), f.x[p], a);
            b = __ffma2_rn(splat2(q1.x), f.x[p], b);
            c = __ffma2_rn(splat2(q2.x), f.x[p], c);
            acc[2 * p] &= __float_as_uint(a.x) | __float_as_uint(b.x) | __float_as_uint(c.x);
            acc[2 * p + 1] &= __float_as_uint(a.y) | __float_as_uint(b.y) | __float_as_uint(c.y);
        }
    }
    uint32_t all = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 2 * RP; ++k) all &= acc[k];
    return all;
}

// V1: scalar FFMAs, one ray at a time (3 register operands per instruction)
template <int RP, int J>
__device__ __forceinline__ uint32_t block_scalar(const Rays<RP>& f, const float4* rec) {
    uint32_t acc[2 * RP];
#pragma unroll
    for (int k = 0; k < 2 * RP; ++k) acc[k] = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q0 = rec[j * 4 + 0], q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2];
#pragma unroll
        for (int k = 0; k < 2 * RP; ++k) {
            const float x = (k & 1) ? f.x[k / 2].y : f.x[k / 2].x, y = (k & 1) ? f.y[k / 2].y : f.y[k / 2].x;
            const float a = fmaf(q0.x, x, fmaf(q0.y, y, q0.z));
            const float b = fmaf(q1.x, x, fmaf(q1.y, y, q1.z));
            const float c = fmaf(q2.x, x, fmaf(q2.y, y, q2.z));
            acc[k] &= __float_as_uint(a) | __float_as_uint(b) | __float_as_uint(c);
        }
    }
    uint32_t all = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 2 * RP; ++k) all &= acc[k];
    return all;
}

// V3: mixed -- first FMA of each chain packed (scalar, pair, scalar: 4 register operands), second scalar (3 operands each)
// V4: mixed the other way round
template <int RP, int J, int V>
__device__ __forceinline__ uint32_t block_mixed(const Rays<RP>& f, const float4* rec) {
    uint32_t acc[2 * RP];
#pragma unroll
    for (int k = 0; k < 2 * RP; ++k) acc[k] = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q0 = rec[j * 4 + 0], q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float2 a, b, c;
            if (V == 3) {
                a = __ffma2_rn(splat2(q0.y), f.y[p], splat2(q0.z));
                b = __ffma2_rn(splat2(q1.y), f.y[p], splat2(q1.z));
                c = __ffma2_rn(splat2(q2.y), f.y[p], splat2(q2.z));
                a.x = fmaf(q0.x, f.x[p].x, a.x); a.y = fmaf(q0.x, f.x[p].y, a.y);
                b.x = fmaf(q1.x, f.x[p].x, b.x); b.y = fmaf(q1.x, f.x[p].y, b.y);
                c.x = fmaf(q2.x, f.x[p].x, c.x); c.y = fmaf(q2.x, f.x[p].y, c.y);
            } else {
                a.x = fmaf(q0.y, f.y[p].x, q0.z); a.y = fmaf(q0.y, f.y[p].y, q0.z);
                b.x = fmaf(q1.y, f.y[p].x, q1.z); b.y = fmaf(q1.y, f.y[p].y, q1.z);
                c.x = fmaf(q2.y, f.y[p].x, q2.z); c.y = fmaf(q2.y, f.y[p].y, q2.z);
                a = __ffma2_rn(splat2(q0.x), f.x[p], a);
                b = __ffma2_rn(splat2(q1.x), f.x[p], b);
                c = __ffma2_rn(splat2(q2.x), f.x[p], c);
            }
            acc[2 * p] &= __float_as_uint(a.x) | __float_as_uint(b.x) | __float_as_uint(c.x);
            acc[2 * p + 1] &= __float_as_uint(a.y) | __float_as_uint(b.y) | __float_as_uint(c.y);
        }
    }
    uint32_t all = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 2 * RP; ++k) all &= acc[k];
    return all;
}

// V5: scalar, two weights; the third weight packed over ray pairs (2 FFMA2 instead of 4 FFMA)
template <int RP, int J>
__device__ __forceinline__ uint32_t block_mixed2(const Rays<RP>& f, const float4* rec) {
    uint32_t acc[2 * RP];
#pragma unroll
    for (int k = 0; k < 2 * RP; ++k) acc[k] = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q0 = rec[j * 4 + 0], q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float2 a, b;
            a.x = fmaf(q0.x, f.x[p].x, fmaf(q0.y, f.y[p].x, q0.z)); a.y = fmaf(q0.x, f.x[p].y, fmaf(q0.y, f.y[p].y, q0.z));
            b.x = fmaf(q1.x, f.x[p].x, fmaf(q1.y, f.y[p].x, q1.z)); b.y = fmaf(q1.x, f.x[p].y, fmaf(q1.y, f.y[p].y, q1.z));
            float2 c = __ffma2_rn(splat2(q2.y), f.y[p], splat2(q2.z));
            c = __ffma2_rn(splat2(q2.x), f.x[p], c);
            acc[2 * p] &= __float_as_uint(a.x) | __float_as_uint(b.x) | __float_as_uint(c.x);
            acc[2 * p + 1] &= __float_as_uint(a.y) | __float_as_uint(b.y) | __float_as_uint(c.y);
        }
    }
    uint32_t all = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 2 * RP; ++k) all &= acc[k];
    return all;
}

// V2: packed FMAs only (no LOP3: results folded into a sum) -- the FMA-side ceiling of the loop
template <int RP, int J>
__device__ __forceinline__ uint32_t block_fma_only(const Rays<RP>& f, const float4* rec, float2& sink) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q0 = rec[j * 4 + 0], q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float2 a = __ffma2_rn(splat2(q0.y), f.y[p], splat2(q0.z));
            float2 b = __ffma2_rn(splat2(q1.y), f.y[p], splat2(q1.z));
            float2 c = __ffma2_rn(splat2(q2.y), f.y[p], splat2(q2.z));
            a = __ffma2_rn(splat2(q0.x), f.x[p], a);
            b = __ffma2_rn(splat2(q1.x), f.x[p], b);
            c = __ffma2_rn(splat2(q2.x), f.x[p], sink);
            sink = __ffma2_rn(a, b, c);   // (one extra packed op per pair-triangle)
        }
    }
    return 0xffffffffu;
}

template <int RP, int J, int V, int MINB>
__global__ void __launch_bounds__(256, MINB) k(const float4* rec_g, float* out, unsigned long long* cyc, float seed) {
    __shared__ float4 tile[kTile * 4];
    for (int i = threadIdx.x; i < kTile * 4; i += blockDim.x) tile[i] = rec_g[i];
    __syncthreads();
    Rays<RP> f;
#pragma unroll
    for (int p = 0; p < RP; ++p) {
        const float t = seed + threadIdx.x * 0.001f + p;
        f.x[p] = make_float2(0.3f + 0.01f * t, 0.31f - 0.01f * t); f.y[p] = make_float2(0.5f + 0.001f * t, 0.49f);
    }
    unsigned hits = 0;
    float2 sink = make_float2(0.f, 0.f);
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll 1
        for (int jb = 0; jb < kTile; jb += J) {
            const uint32_t all = (V == 0) ? block_packed<RP, J>(f, tile + jb * 4) : (V == 1) ? block_scalar<RP, J>(f, tile + jb * 4)
                               : (V == 3 || V == 4) ? block_mixed<RP, J, V>(f, tile + jb * 4) : (V == 5) ? block_mixed2<RP, J>(f, tile + jb * 4) : block_fma_only<RP, J>(f, tile + jb * 4, sink);
            if ((int)all >= 0) { ++hits; f.x[0].x += 1e-6f * hits; }   // rare side effect so nothing is optimised away
        }
    }
    unsigned long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)hits + sink.x + sink.y;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int RP, int J, int V, int MINB>
void run(const char* name, const float4* rec, float* out, unsigned long long* cyc) {
    k<RP, J, V, MINB><<<148 * MINB, 256>>>(rec, out, cyc, 1.0f);   // warm-up
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<RP, J, V, MINB><<<148 * MINB, 256>>>(rec, out, cyc, 1.0f);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)ITERS * kTile * RP * (MINB * 2));   // warps per SMSP = MINB * 8 / 4
    const double tests = 148.0 * MINB * 256 * 2 * RP * (double)ITERS * kTile;
    printf("%-40s %7.2f cycles per (pair,tri) per SMSP; %.3f ms -> %.3e tests/s  (%s)\n", name, per, ms, tests / (ms * 1e-3), cudaGetErrorString(e));
}

int main() {
    float4* rec; float* out; unsigned long long* cyc;
    cudaMalloc(&rec, kTile * 4 * sizeof(float4)); cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&cyc, 8);
    float4 h[kTile * 4];
    for (int i = 0; i < kTile; ++i) {   // weights that are negative for every ray: no candidates
        h[4 * i] = make_float4(0.1f + 0.001f * i, 0.7f, -5.f, 0.f);
        h[4 * i + 1] = make_float4(1.f, 0.5f, -7.f, 0.f);
        h[4 * i + 2] = make_float4(-1.f, 0.75f, -3.f, 0.f);
        h[4 * i + 3] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cudaMemcpy(rec, h, sizeof(h), cudaMemcpyHostToDevice);
    run<2, 8, 0, 2>("packed rp2 j8 minb2 (shipped)", rec, out, cyc);
    run<3, 8, 0, 2>("packed rp3 j8 minb2", rec, out, cyc);
    run<4, 8, 0, 2>("packed rp4 j8 minb2", rec, out, cyc);
    run<4, 4, 0, 2>("packed rp4 j4 minb2", rec, out, cyc);
    run<4, 8, 0, 1>("packed rp4 j8 minb1", rec, out, cyc);
    run<6, 4, 0, 1>("packed rp6 j4 minb1", rec, out, cyc);
    run<2, 8, 1, 2>("scalar rp2 j8 minb2", rec, out, cyc);
    run<2, 8, 2, 2>("packed FMA only rp2 j8 minb2 (7 ops)", rec, out, cyc);
    run<4, 8, 1, 2>("scalar rp4 j8 minb2", rec, out, cyc);
    run<4, 4, 1, 2>("scalar rp4 j4 minb2", rec, out, cyc);
    run<3, 8, 1, 2>("scalar rp3 j8 minb2", rec, out, cyc);
    run<2, 8, 3, 2>("mixed(packed first) rp2 j8", rec, out, cyc);
    run<3, 8, 3, 2>("mixed(packed first) rp3 j8", rec, out, cyc);
    run<4, 4, 3, 2>("mixed(packed first) rp4 j4", rec, out, cyc);
    run<4, 8, 3, 2>("mixed(packed first) rp4 j8", rec, out, cyc);
    run<2, 8, 4, 2>("mixed(packed second) rp2 j8", rec, out, cyc);
    run<3, 8, 4, 2>("mixed(packed second) rp3 j8", rec, out, cyc);
    run<4, 4, 4, 2>("mixed(packed second) rp4 j4", rec, out, cyc);
    run<4, 8, 4, 2>("mixed(packed second) rp4 j8", rec, out, cyc);
    run<2, 8, 5, 2>("2 scalar weights + 1 packed rp2 j8", rec, out, cyc);
    run<3, 8, 5, 2>("2 scalar weights + 1 packed rp3 j8", rec, out, cyc);
    run<4, 4, 5, 2>("2 scalar weights + 1 packed rp4 j4", rec, out, cyc);
    run<4, 8, 5, 2>("2 scalar weights + 1 packed rp4 j8", rec, out, cyc);
    run<4, 4, 1, 1>("scalar rp4 j4 minb1", rec, out, cyc);
    run<4, 8, 1, 1>("scalar rp4 j8 minb1", rec, out, cyc);
    return 0;
}
