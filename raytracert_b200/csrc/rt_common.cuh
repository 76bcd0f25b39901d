// rt_common.cuh -- device-side building blocks of librt_b200 (sm_100a only).
//
//  * "exact" arithmetic: the reference's float expressions restated op for op with the round-to-nearest
//    intrinsics (__fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn).  nvcc never contracts these into
//    FMAs, and IEEE binary32 +,-,*,/,sqrt are correctly rounded on both x86-64 SSE2 and sm_100a, so these
//    functions return the SAME BITS as the reference's CPU code (build without -use_fast_math / -ftz).
//  * mbarrier + 1-D TMA bulk-copy wrappers (cp.async.bulk ... mbarrier::complete_tx::bytes) used to stage
//    triangle tiles into shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace rt {

// ------------------------------------------------------------------------------------------------
// exact (reference-order) 3-vector maths -- Vec3D.h of the reference
// ------------------------------------------------------------------------------------------------
struct v3 { float x, y, z; };

__device__ __forceinline__ v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 mk3(const float4& f) { return mk3(f.x, f.y, f.z); }
__device__ __forceinline__ v3 e_add(v3 a, v3 b) { return mk3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }   // Vec3D.h:24-26
__device__ __forceinline__ v3 e_sub(v3 a, v3 b) { return mk3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }   // Vec3D.h:28-30
__device__ __forceinline__ v3 e_neg(v3 a) { return mk3(-a.x, -a.y, -a.z); }                                                          // Vec3D.h:32-34
__device__ __forceinline__ v3 e_scale(v3 a, float s) { return mk3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }      // Vec3D.h:12-18
__device__ __forceinline__ v3 e_mul(v3 a, v3 b) { return mk3(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z)); }   // Vec3D.h:20-22
__device__ __forceinline__ v3 e_div(v3 a, float s) { return mk3(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)); }        // Vec3D.h:36-38
__device__ __forceinline__ float e_dot(v3 a, v3 b) {                                                                                  // Vec3D.h:192-194
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ v3 e_cross(v3 a, v3 b) {                                                                                   // Vec3D.h:185-191
    return mk3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)),
               __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
               __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// (float)sqrt((double)x) == correctly rounded float sqrt (53 >= 2*24+2 bits), Vec3D.h:135-140
__device__ __forceinline__ float e_length(v3 a) { return __fsqrt_rn(e_dot(a, a)); }
__device__ __forceinline__ v3 e_normalize(v3 a) {                                                                                     // Vec3D.h:142-151
    float len = e_length(a);
    if (len == 0.0f) return a;
    float rez = __fdiv_rn(1.0f, len);
    return e_scale(a, rez);
}
__device__ __forceinline__ float e_distance(v3 a, v3 b) { return e_length(e_sub(a, b)); }                                             // Vec3D.h:199-202
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }

// rayIntersectTriangle, raytracing.cpp:99-154.  Returns true on a hit and writes the intersection point.
__device__ __forceinline__ bool exact_ray_triangle(v3 R0, v3 R1, v3 T0, v3 T1, v3 T2, v3& I_out) {
    const float SMALL_NUM = 0.00001f;
    v3 u = e_sub(T1, T0);
    v3 v = e_sub(T2, T0);
    v3 n = e_cross(u, v);
    if (n.x == 0.0f && n.y == 0.0f && n.z == 0.0f) return false;   // :109
    v3 dir = e_sub(R1, R0);
    v3 w0 = e_sub(R0, T0);
    float b = e_dot(n, dir);
    float a = -e_dot(n, w0);
    if (fabsf(b) < SMALL_NUM) return false;                         // :115
    float r = __fdiv_rn(a, b);
    if (r < 0.0f) return false;                                     // :125
    v3 I = e_add(R0, e_scale(dir, r));                              // :130
    float uu = e_dot(u, u), uv = e_dot(u, v), vv = e_dot(v, v);
    v3 w = e_sub(I, T0);
    float wu = e_dot(w, u), wv = e_dot(w, v);
    float D = __fsub_rn(__fmul_rn(uv, uv), __fmul_rn(uu, vv));
    float s = __fdiv_rn(__fsub_rn(__fmul_rn(uv, wv), __fmul_rn(vv, wu)), D);
    if (s < 0.0f || s > 1.0f) return false;                         // :145 (NaN passes)
    float t = __fdiv_rn(__fsub_rn(__fmul_rn(uv, wu), __fmul_rn(uu, wv)), D);
    if (t < 0.0f || __fadd_rn(s, t) > 1.0f) return false;           // :149
    I_out = I;
    return true;
}

// Sphere primitive -- this repo's own semantics (Sphere.h of the reference is orphaned);
// identical to oracle/rt_oracle.c:ray_intersect_sphere.
__device__ __forceinline__ bool exact_ray_sphere(v3 R0, v3 R1, v3 C, float radius, v3& I_out) {
    v3 d = e_normalize(e_sub(R1, R0));
    v3 oc = e_sub(R0, C);
    float bq = e_dot(oc, d);
    float cq = __fsub_rn(e_dot(oc, oc), __fmul_rn(radius, radius));
    float disc = __fsub_rn(__fmul_rn(bq, bq), cq);
    if (disc < 0.0f) return false;
    float sq = __fsqrt_rn(disc);
    float t = __fsub_rn(-bq, sq);
    if (!(t > 1e-4f)) t = __fadd_rn(-bq, sq);
    if (!(t > 1e-4f)) return false;
    I_out = e_add(R0, e_scale(d, t));
    return true;
}

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk copy (PTX)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP). 16-byte aligned.
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }

// Warp-aggregated append: every lane of the (fully converged) warp calls this; lanes with pred push `value`.
__device__ __forceinline__ void warp_append(bool pred, uint32_t value, uint32_t* __restrict__ queue, uint32_t* __restrict__ counter) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) queue[base + __popc(m & ((1u << lane) - 1u))] = value;
}

}  // namespace rt
