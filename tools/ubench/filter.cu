// Micro-benchmark (exploration, not product): instruction ordering of the filter loop vs register-file
// operand bandwidth on sm_100a.  One 128-triangle tile of records sits in shared memory and is scanned
// ITERS times by every warp; prints SM cycles per (ray pair, triangle).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo filter.cu -o filter
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kTile = 128;
#define ITERS 256

__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int RP>
struct Rays { float2 ox[RP], oy[RP], oz[RP], dx[RP], dy[RP], dz[RP]; uint32_t rhi[2 * RP]; };

// V0: the shipped order -- for each triangle, for each ray pair, the whole chain
template <int RP, int J>
__device__ __forceinline__ bool block_v0(const Rays<RP>& f, const float4* rec) {
    bool any = false;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q0 = rec[j * 4 + 0], q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2], q3 = rec[j * 4 + 3];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float2 b = __fmul2_rn(splat2(q0.x), f.dx[p]);
            b = __ffma2_rn(splat2(q0.y), f.dy[p], b);
            b = __ffma2_rn(splat2(q0.z), f.dz[p], b);
            float2 a = __ffma2_rn(splat2(q0.x), f.ox[p], splat2(q0.w));
            a = __ffma2_rn(splat2(q0.y), f.oy[p], a);
            a = __ffma2_rn(splat2(q0.z), f.oz[p], a);
            const float2 rc = make_float2(rcp_approx(-b.x), rcp_approx(-b.y));
            const float2 r = __fmul2_rn(a, rc);
            const float2 ix = __ffma2_rn(r, f.dx[p], f.ox[p]);
            const float2 iy = __ffma2_rn(r, f.dy[p], f.oy[p]);
            const float2 iz = __ffma2_rn(r, f.dz[p], f.oz[p]);
            float2 s = __ffma2_rn(splat2(q1.x), ix, splat2(q1.w));
            s = __ffma2_rn(splat2(q1.y), iy, s);
            s = __ffma2_rn(splat2(q1.z), iz, s);
            float2 t = __ffma2_rn(splat2(q2.x), ix, splat2(q2.w));
            t = __ffma2_rn(splat2(q2.y), iy, t);
            t = __ffma2_rn(splat2(q2.z), iz, t);
            float2 q = __fadd2_rn(splat2(q3.x), make_float2(-s.x, -s.y));
            q = __fadd2_rn(q, make_float2(-t.x, -t.y));
            const float m0 = fminf(fminf(s.x, t.x), q.x), m1 = fminf(fminf(s.y, t.y), q.y);
            const float2 e = __fmul2_rn(splat2(q3.y), rc);
            const bool c0 = (!(m0 < -fabsf(e.x)) && (__float_as_uint(r.x) < f.rhi[2 * p])) || (fabsf(b.x) < q3.z);
            const bool c1 = (!(m1 < -fabsf(e.y)) && (__float_as_uint(r.y) < f.rhi[2 * p + 1])) || (fabsf(b.y) < q3.z);
            any = any || c0 || c1;
        }
    }
    return any;
}

// V1: phase order -- for a fixed ray pair, every phase runs over the J triangles, so consecutive packed FMAs
// share the ray operand (operand reuse cache) and differ in the triangle scalar + accumulator
template <int RP, int J>
__device__ __forceinline__ bool block_v1(const Rays<RP>& f, const float4* rec) {
    bool any = false;
#pragma unroll
    for (int p = 0; p < RP; ++p) {
        float4 q0[J];
        float2 b[J], a[J], r[J], rc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) q0[j] = rec[j * 4 + 0];
#pragma unroll
        for (int j = 0; j < J; ++j) b[j] = __fmul2_rn(splat2(q0[j].x), f.dx[p]);
#pragma unroll
        for (int j = 0; j < J; ++j) b[j] = __ffma2_rn(splat2(q0[j].y), f.dy[p], b[j]);
#pragma unroll
        for (int j = 0; j < J; ++j) b[j] = __ffma2_rn(splat2(q0[j].z), f.dz[p], b[j]);
#pragma unroll
        for (int j = 0; j < J; ++j) a[j] = __ffma2_rn(splat2(q0[j].x), f.ox[p], splat2(q0[j].w));
#pragma unroll
        for (int j = 0; j < J; ++j) a[j] = __ffma2_rn(splat2(q0[j].y), f.oy[p], a[j]);
#pragma unroll
        for (int j = 0; j < J; ++j) a[j] = __ffma2_rn(splat2(q0[j].z), f.oz[p], a[j]);
#pragma unroll
        for (int j = 0; j < J; ++j) { rc[j] = make_float2(rcp_approx(-b[j].x), rcp_approx(-b[j].y)); r[j] = __fmul2_rn(a[j], rc[j]); }
        float2 ix[J], iy[J], iz[J], s[J], t[J];
#pragma unroll
        for (int j = 0; j < J; ++j) ix[j] = __ffma2_rn(r[j], f.dx[p], f.ox[p]);
#pragma unroll
        for (int j = 0; j < J; ++j) iy[j] = __ffma2_rn(r[j], f.dy[p], f.oy[p]);
#pragma unroll
        for (int j = 0; j < J; ++j) iz[j] = __ffma2_rn(r[j], f.dz[p], f.oz[p]);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2];
            s[j] = __ffma2_rn(splat2(q1.x), ix[j], splat2(q1.w));
            t[j] = __ffma2_rn(splat2(q2.x), ix[j], splat2(q2.w));
            s[j] = __ffma2_rn(splat2(q1.y), iy[j], s[j]);
            t[j] = __ffma2_rn(splat2(q2.y), iy[j], t[j]);
            s[j] = __ffma2_rn(splat2(q1.z), iz[j], s[j]);
            t[j] = __ffma2_rn(splat2(q2.z), iz[j], t[j]);
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 q3 = rec[j * 4 + 3];
            float2 q = __fadd2_rn(splat2(q3.x), make_float2(-s[j].x, -s[j].y));
            q = __fadd2_rn(q, make_float2(-t[j].x, -t[j].y));
            const float m0 = fminf(fminf(s[j].x, t[j].x), q.x), m1 = fminf(fminf(s[j].y, t[j].y), q.y);
            const float2 e = __fmul2_rn(splat2(q3.y), rc[j]);
            const bool c0 = (!(m0 < -fabsf(e.x)) && (__float_as_uint(r[j].x) < f.rhi[2 * p])) || (fabsf(b[j].x) < q3.z);
            const bool c1 = (!(m1 < -fabsf(e.y)) && (__float_as_uint(r[j].y) < f.rhi[2 * p + 1])) || (fabsf(b[j].y) < q3.z);
            any = any || c0 || c1;
        }
    }
    return any;
}


// V2: the 19 packed ops + MUFU, no compare logic (results folded into a running sum)
// V3: V2 without the MUFU (r = a*b)      V4: V0 without the |cos| clause     V5: V0 with the clause but no min3/e (s only)
template <int RP, int J, int V>
__device__ __forceinline__ bool block_vx(const Rays<RP>& f, const float4* rec, float2& sink) {
    bool any = false;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q0 = rec[j * 4 + 0], q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2], q3 = rec[j * 4 + 3];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float2 b = __fmul2_rn(splat2(q0.x), f.dx[p]);
            b = __ffma2_rn(splat2(q0.y), f.dy[p], b);
            b = __ffma2_rn(splat2(q0.z), f.dz[p], b);
            float2 a = __ffma2_rn(splat2(q0.x), f.ox[p], splat2(q0.w));
            a = __ffma2_rn(splat2(q0.y), f.oy[p], a);
            a = __ffma2_rn(splat2(q0.z), f.oz[p], a);
            const float2 rc = (V == 3) ? b : make_float2(rcp_approx(-b.x), rcp_approx(-b.y));
            const float2 r = __fmul2_rn(a, rc);
            const float2 ix = __ffma2_rn(r, f.dx[p], f.ox[p]);
            const float2 iy = __ffma2_rn(r, f.dy[p], f.oy[p]);
            const float2 iz = __ffma2_rn(r, f.dz[p], f.oz[p]);
            float2 s = __ffma2_rn(splat2(q1.x), ix, splat2(q1.w));
            s = __ffma2_rn(splat2(q1.y), iy, s);
            s = __ffma2_rn(splat2(q1.z), iz, s);
            float2 t = __ffma2_rn(splat2(q2.x), ix, splat2(q2.w));
            t = __ffma2_rn(splat2(q2.y), iy, t);
            t = __ffma2_rn(splat2(q2.z), iz, t);
            float2 q = __fadd2_rn(splat2(q3.x), make_float2(-s.x, -s.y));
            q = __fadd2_rn(q, make_float2(-t.x, -t.y));
            const float2 e = __fmul2_rn(splat2(q3.y), rc);
            if (V == 2 || V == 3) { sink = __ffma2_rn(q, e, sink); }
            if (V == 4) {
                const float m0 = fminf(fminf(s.x, t.x), q.x), m1 = fminf(fminf(s.y, t.y), q.y);
                const bool c0 = (!(m0 < -fabsf(e.x)) && (__float_as_uint(r.x) < f.rhi[2 * p]));
                const bool c1 = (!(m1 < -fabsf(e.y)) && (__float_as_uint(r.y) < f.rhi[2 * p + 1]));
                any = any || c0 || c1;
            }
            if (V == 5) {
                const bool c0 = (!(q.x < -fabsf(e.x)) && (__float_as_uint(r.x) < f.rhi[2 * p])) || (fabsf(b.x) < q3.z);
                const bool c1 = (!(q.y < -fabsf(e.y)) && (__float_as_uint(r.y) < f.rhi[2 * p + 1])) || (fabsf(b.y) < q3.z);
                any = any || c0 || c1;
            }
        }
    }
    return any;
}

// V6: 2-D projected barycentrics (dominant-axis classes): I needs 2 components, s/t are 2-term -> 16 packed ops
template <int RP, int J>
__device__ __forceinline__ bool block_v6(const Rays<RP>& f, const float4* rec) {
    bool any = false;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q0 = rec[j * 4 + 0], q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2], q3 = rec[j * 4 + 3];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float2 b = __fmul2_rn(splat2(q0.x), f.dx[p]);
            b = __ffma2_rn(splat2(q0.y), f.dy[p], b);
            b = __ffma2_rn(splat2(q0.z), f.dz[p], b);
            float2 a = __ffma2_rn(splat2(q0.x), f.ox[p], splat2(q0.w));
            a = __ffma2_rn(splat2(q0.y), f.oy[p], a);
            a = __ffma2_rn(splat2(q0.z), f.oz[p], a);
            const float2 rc = make_float2(rcp_approx(-b.x), rcp_approx(-b.y));
            const float2 r = __fmul2_rn(a, rc);
            const float2 iu = __ffma2_rn(r, f.dx[p], f.ox[p]);
            const float2 iv = __ffma2_rn(r, f.dy[p], f.oy[p]);
            float2 s = __ffma2_rn(splat2(q1.x), iu, splat2(q1.z));
            s = __ffma2_rn(splat2(q1.y), iv, s);
            float2 t = __ffma2_rn(splat2(q2.x), iu, splat2(q2.z));
            t = __ffma2_rn(splat2(q2.y), iv, t);
            float2 q = __fadd2_rn(splat2(q1.w), make_float2(-s.x, -s.y));
            q = __fadd2_rn(q, make_float2(-t.x, -t.y));
            const float m0 = fminf(fminf(s.x, t.x), q.x), m1 = fminf(fminf(s.y, t.y), q.y);
            const float2 e = __fmul2_rn(splat2(q2.w), rc);
            const bool c0 = (!(m0 < -fabsf(e.x)) && (__float_as_uint(r.x) < f.rhi[2 * p])) || (fabsf(b.x) < q3.x);
            const bool c1 = (!(m1 < -fabsf(e.y)) && (__float_as_uint(r.y) < f.rhi[2 * p + 1])) || (fabsf(b.y) < q3.x);
            any = any || c0 || c1;
        }
    }
    return any;
}

// V7: V6 with scalar FFMAs (3 register operands each) and without the grazing clause
// V8: V6 packed, without the grazing clause (the shipped clause-free kernels)
template <int RP, int J, int V>
__device__ __forceinline__ bool block_v78(const Rays<RP>& f, const float4* rec) {
    bool any = false;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q0 = rec[j * 4 + 0], q1 = rec[j * 4 + 1], q2 = rec[j * 4 + 2];
        if (V == 8) {
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                float2 b = __fmul2_rn(splat2(q0.x), f.dx[p]);
                b = __ffma2_rn(splat2(q0.y), f.dy[p], b);
                b = __ffma2_rn(splat2(q0.z), f.dz[p], b);
                float2 a = __ffma2_rn(splat2(q0.x), f.ox[p], splat2(q0.w));
                a = __ffma2_rn(splat2(q0.y), f.oy[p], a);
                a = __ffma2_rn(splat2(q0.z), f.oz[p], a);
                const float2 rc = make_float2(rcp_approx(-b.x), rcp_approx(-b.y));
                const float2 r = __fmul2_rn(a, rc);
                const float2 iu = __ffma2_rn(r, f.dx[p], f.ox[p]);
                const float2 iv = __ffma2_rn(r, f.dy[p], f.oy[p]);
                float2 s = __ffma2_rn(splat2(q1.x), iu, splat2(q1.z));
                s = __ffma2_rn(splat2(q1.y), iv, s);
                float2 t = __ffma2_rn(splat2(q2.x), iu, splat2(q2.z));
                t = __ffma2_rn(splat2(q2.y), iv, t);
                float2 q = __fadd2_rn(splat2(q1.w), make_float2(-s.x, -s.y));
                q = __fadd2_rn(q, make_float2(-t.x, -t.y));
                const float m0 = fminf(fminf(s.x, t.x), q.x), m1 = fminf(fminf(s.y, t.y), q.y);
                const float2 e = __fmul2_rn(splat2(q2.w), rc);
                const bool c0 = (!(m0 < -fabsf(e.x)) && (__float_as_uint(r.x) < f.rhi[2 * p]));
                const bool c1 = (!(m1 < -fabsf(e.y)) && (__float_as_uint(r.y) < f.rhi[2 * p + 1]));
                any = any || c0 || c1;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 2 * RP; ++k) {
                const int p = k / 2;
                const float dx = (k & 1) ? f.dx[p].y : f.dx[p].x, dy = (k & 1) ? f.dy[p].y : f.dy[p].x, dz = (k & 1) ? f.dz[p].y : f.dz[p].x;
                const float ox = (k & 1) ? f.ox[p].y : f.ox[p].x, oy = (k & 1) ? f.oy[p].y : f.oy[p].x, oz = (k & 1) ? f.oz[p].y : f.oz[p].x;
                const float b = fmaf(q0.z, dz, fmaf(q0.y, dy, q0.x * dx));
                const float a = fmaf(q0.z, oz, fmaf(q0.y, oy, fmaf(q0.x, ox, q0.w)));
                const float rc = rcp_approx(-b);
                const float r = a * rc;
                const float iu = fmaf(r, dx, ox), iv = fmaf(r, dy, oy);
                const float s = fmaf(q1.y, iv, fmaf(q1.x, iu, q1.z));
                const float t = fmaf(q2.y, iv, fmaf(q2.x, iu, q2.z));
                const float q = (q1.w - s) - t;
                const float m = fminf(fminf(s, t), q);
                const float e = q2.w * rc;
                any = any || (!(m < -fabsf(e)) && (__float_as_uint(r) < f.rhi[k]));
            }
        }
    }
    return any;
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
        f.ox[p] = make_float2(t, t + 1); f.oy[p] = make_float2(2 * t, t - 1); f.oz[p] = make_float2(-t, 3 - t);
        f.dx[p] = make_float2(0.3f + 0.01f * t, 0.31f - 0.01f * t); f.dy[p] = make_float2(0.5f + 0.001f * t, 0.49f); f.dz[p] = make_float2(-0.81f, -0.8f + 0.002f * t);
        f.rhi[2 * p] = 0x7f7fffffu; f.rhi[2 * p + 1] = 0x7f7fffffu;
    }
    unsigned hits = 0;
    float2 sink = make_float2(0.f, 0.f);
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll 1
        for (int jb = 0; jb < kTile; jb += J) {
            const bool any = (V == 0) ? block_v0<RP, J>(f, tile + jb * 4) : (V == 1) ? block_v1<RP, J>(f, tile + jb * 4) : (V == 6) ? block_v6<RP, J>(f, tile + jb * 4)
                           : (V == 7 || V == 8) ? block_v78<RP, J, V>(f, tile + jb * 4) : block_vx<RP, J, V>(f, tile + jb * 4, sink);
            if (any) { ++hits; f.rhi[0] ^= hits; }   // rare side effect so nothing is optimised away
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
    // warps per SMSP = MINB * 8 / 4; cycles per (pair, triangle) per SMSP = wall / (ITERS * kTile * RP * warps_per_smsp)
    const double per = (double)h / ((double)ITERS * kTile * RP * (MINB * 2));
    const double tests = 148.0 * MINB * 256 * 2 * RP * (double)ITERS * kTile;
    printf("%-34s %8.2f clock64 cycles per (pair,tri) per SMSP; event %.3f ms -> %.3e tests/s  (%s)\n", name, per, ms, tests / (ms * 1e-3), cudaGetErrorString(e));
}

int main() {
    float4* rec; float* out; unsigned long long* cyc;
    cudaMalloc(&rec, kTile * 4 * sizeof(float4)); cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&cyc, 8);
    float4 h[kTile * 4];
    for (int i = 0; i < kTile; ++i) {   // planes far from the rays: no candidates
        h[4 * i] = make_float4(0.1f + 0.001f * i, 0.7f, 0.7f, 50.f + i);
        h[4 * i + 1] = make_float4(1.f, 0.5f, 0.25f, 0.125f);
        h[4 * i + 2] = make_float4(-1.f, 0.75f, 0.5f, 0.25f);
        h[4 * i + 3] = make_float4(1.001f, -1e-6f, 1e-5f, 0.f);
    }
    cudaMemcpy(rec, h, sizeof(h), cudaMemcpyHostToDevice);
    run<2, 8, 6, 2>("V6 2-D projected (16 packed)", rec, out, cyc);
    run<2, 8, 8, 2>("V8 V6 clause-free packed rp2 j8", rec, out, cyc);
    run<2, 8, 7, 2>("V7 clause-free scalar rp2 j8", rec, out, cyc);
    run<1, 16, 7, 2>("V7 clause-free scalar rp1 j16 minb2", rec, out, cyc);
    run<1, 16, 7, 4>("V7 clause-free scalar rp1 j16 minb4", rec, out, cyc);
    run<2, 8, 7, 3>("V7 clause-free scalar rp2 j8 minb3", rec, out, cyc);
    run<3, 4, 7, 2>("V7 clause-free scalar rp3 j4", rec, out, cyc);
    run<3, 4, 8, 2>("V8 clause-free packed rp3 j4", rec, out, cyc);
    return 0;
}
