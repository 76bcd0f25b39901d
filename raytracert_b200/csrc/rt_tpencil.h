// rt_tpencil.h -- "thread pencils": the R rays of ONE THREAD share a common point.
//
// Every facet is a plane: the continuation rays (reflection(), raytracing.cpp:277-285) of the primary hits on ONE triangle T
// leave the mirror image E*_T of the eye about T's plane -- a pencil, like the reflection pencils of rt_pencil.h, but one per
// triangle.  A set of pencil records per triangle is out of the question (a set costs ~550 rays' worth of savings; a
// triangle of the headline frame carries 80-240 level-1 rays).  But the pencil weights are AFFINE in the common point,
//     (v1 - E) x (v2 - E) = v1 x v2 + E x (v1 - v2),
// so a thread whose rays share E can build its three weight vectors per triangle from an E-independent record in 18 FMAs
// (+ 3 for the side of the plane E lies on, + 9 sign flips) and then spend 9 FMAs + 1.5 LOP3 per ray: scalar code, ~30 issue
// slots per (ray pair, triangle) with 8 rays per thread against the generic filter's 52 cycles.
//
// Shared by the CUDA library (k_build_trec, k_tp_route, k_trace_tp) and the CPU soundness replay (tests/pencil_check.cpp,
// mode 4): everything the filter evaluates is written with explicit fmaf() / single IEEE operations in a fixed order, so
// the CPU replay performs the same arithmetic as the kernel.
//
// Soundness (on top of rt_pencil.h's argument, whose constants E0, E1, K_r, s_lam are reused):
//   * coordinates are taken relative to the centre of the scene box (smaller magnitudes in v1 x v2);
//   * a ray is only ever put on its triangle's pencil after tp_accepts() has CHECKED, in float, on the ray as built, that its
//     line passes within delta/2 of the E* the scan will use (the same device function computes it both times), that E* is
//     at least lam_min behind the origin and inside the magnitude bound Emax;
//   * oriented weights: X(E) = X0' + E x dX with X0' = X0 + E0*n folded in at build time (E-independent); the constant s covers
//     1.5*E1p*|n| and the rounding of the on-the-fly construction and of the 3-FMA dot product (<= 16u*(|X0'|_1 + 2*Emax*|dX|_1));
//   * side of the plane: det = (v0 - E).n = c0 - E.n; a valid pair with |cos| >= cos_g has |det|/|n| >= lam_min*cos_g - delta
//     >= 1.5*delta_eff, far above det's rounding, so its sign is right whenever it matters;
//   * pairs with |cos| < cos_g (cos_g = max(1.05e-5, 2.5*delta_eff/lam_min_thread, 5*theta)) can only exist for a triangle whose plane
//     passes within lam_max*cos_g + 2*delta_eff of E* ("near"): the hot loop makes a near triangle a candidate for all the
//     thread's rays, and the cold path's full test adds the grazing clause |sigma| < cos_g*|n| for it;
//   * distance clauses (cold path only): |det| - KrN < sigma_s*lam_hi and |det| + KrN > sigma_s*lam_lo with lam_hi = lam_O +
//     nearest + slack, lam_lo = lam_O - slack (the hit must lie in front of the ray's origin: E* is BEHIND the reflector).
#pragma once
#include "rt_pencil.h"

namespace rt {

constexpr int kTpVec = 6;   // float4 per thread-pencil record (96 B)

struct TpSetup {
    double center[3];   // coordinates are relative to this point
    double M;           // bound on |coordinate| (scene, ray origins), as everywhere
    double Mc;          // bound on |centred coordinate| of the vertices
    double Emax;        // bound on |centred coordinate| of an accepted E*
    double delta;       // accepted rays pass within delta of their E* (the check enforces delta/2; the rest is its own rounding)
    double delta_eff;   // delta + rounding of det in distance units
    double theta;       // angular error of the unit direction w
    double lam_max;     // >= |X - E*| for scene points X
    double lam_min;     // accepted rays have |O - E*| >= lam_min
    // float copies for the device / the replay
    float eye[3], centerf[3];
    float half_delta2, lam_min2, emax;
    float lam_slack;    // ray-side slack of the distance clauses
    float cg_num;       // 2.5 * delta_eff      (cos_g = max(cg_floor, cg_num / lam_min_thread))
    float cg_floor;     // max(1.05e-5, 5 * theta)
    float near_a;       // lam_max * 1.001      (near iff |det| < (near_a * cos_g + near_b) * nlen)
    float near_b;       // 2 * delta_eff * 1.001
};

// Launch constants.  eye: the camera pencil's centre (double); delta_cam: its delta; box: everything a record can make a hit of.
inline bool tp_setup(const double eye[3], double delta_cam, double M_scene, const float box_lo[3], const float box_hi[3], TpSetup& S) {
    if (!(M_scene < 1e18)) return false;
    double R2 = 0.0, ec2 = 0.0, Mc = 0.0;
    for (int k = 0; k < 3; ++k) {
        if (!(box_lo[k] <= box_hi[k])) return false;
        S.center[k] = 0.5 * ((double)box_lo[k] + box_hi[k]);
        const double h = 0.5 * ((double)box_hi[k] - box_lo[k]) + 0.2;
        R2 += h * h; Mc = fmax(Mc, h);
        ec2 += (eye[k] - S.center[k]) * (eye[k] - S.center[k]);
        S.eye[k] = (float)eye[k]; S.centerf[k] = (float)S.center[k];
        if (!isfinite(eye[k])) return false;
    }
    const double R = sqrt(R2), ec = sqrt(ec2);
    S.M = fmax(M_scene, 1e-3);
    S.Mc = Mc * 1.001 + 1e-6 * S.M;
    // a mirror image of the eye about a plane that meets the box is as far from any point of that plane as the eye is
    S.Emax = (ec + 2.0 * R) * 1.01 + 1e-3 * S.M;
    S.lam_max = (ec + 3.0 * R) * 1.01 + 1e-3 * S.M;
    S.theta = 8.0 * kPencilU;
    // the check passes rays within delta/2 of E*; its own rounding (|(O - E) x d| to 8u|O - E||d|: 8u*lam_max) stays inside the other half
    S.delta = 4.0 * delta_cam + 24.0 * kPencilU * fmax(S.M, S.lam_max);
    // det = c0 - E.n in float: three FMAs on rounded c0, n, E -> <= 6u(|v0| + |E|)|n|, i.e. 6u*sqrt(3)*(Mc + Emax) in distance units
    S.delta_eff = S.delta + 6.0 * kPencilU * 1.7320508 * (S.Mc + S.Emax);
    S.lam_min = 1e-2 * S.M;
    if (!(S.delta_eff <= 1e-3 * S.M)) return false;
    S.half_delta2 = (float)(0.25 * S.delta * S.delta * 0.999);
    S.lam_min2 = (float)(S.lam_min * S.lam_min * 1.001);
    S.emax = (float)(S.Emax * 0.999);
    S.lam_slack = (float)((128.0 * kPencilU * S.M + 3.0 * S.delta_eff + 8.0 * kPencilU * S.lam_max) * 1.0001);
    S.cg_num = (float)(2.5 * S.delta_eff * 1.001);
    S.cg_floor = (float)fmax(kPencilCosMin, 5.0 * S.theta);
    S.near_a = (float)(S.lam_max * 1.001);
    S.near_b = (float)(2.0 * S.delta_eff * 1.001);
    return true;
}

// Record of one triangle (24 floats), E-independent.  Returns false (and a "never" record) for a degenerate triangle.
//   q0 = ( A0'xyz, s )     A0' = v1 x v2 + E0*n   (weight of v0)          s: constant of the three FMA chains
//   q1 = ( B0'xyz, id )    B0' = v2 x v0 + E0*n                            id / nv: filled in by the caller
//   q2 = ( C0'xyz, nv )    C0' = v0 x v1 + E0*n
//   q3 = ( da xyz, c0 )    da = v1 - v2,  c0 = v0.n
//   q4 = ( db xyz, nlen )  db = v2 - v0,  nlen >= |n|          (dc = v0 - v1 = -(da + db))
//   q5 = ( n xyz, KrN )    n = (v1 - v0) x (v2 - v0),  KrN: guard band of the distance clauses in det units
RT_HD void tp_never(float q[24]) {
    for (int i = 0; i < 24; ++i) q[i] = 0.0f;
    q[3] = -1.0f;
}
RT_HD bool tp_record(const float A[3], const float B[3], const float C[3], double E0, double E1, const TpSetup& S, float q[24]) {
    double v0[3], v1[3], v2[3], da[3], db[3], dc[3], u[3], v[3];
    for (int k = 0; k < 3; ++k) {
        v0[k] = (double)A[k] - S.center[k]; v1[k] = (double)B[k] - S.center[k]; v2[k] = (double)C[k] - S.center[k];
        da[k] = v1[k] - v2[k]; db[k] = v2[k] - v0[k]; dc[k] = v0[k] - v1[k];
        u[k] = v1[k] - v0[k]; v[k] = v2[k] - v0[k];
    }
    const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (!(nn > 0.0) || !isfinite(nn)) { tp_never(q); return false; }
    const double X0[3][3] = {{v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]},
                             {v2[1] * v0[2] - v2[2] * v0[1], v2[2] * v0[0] - v2[0] * v0[2], v2[0] * v0[1] - v2[1] * v0[0]},
                             {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]}};
    const double* dX[3] = {da, db, dc};
    const double l12 = sqrt(da[0] * da[0] + da[1] * da[1] + da[2] * da[2]), l20 = sqrt(db[0] * db[0] + db[1] * db[1] + db[2] * db[2]);
    const double l01 = sqrt(dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2]);
    const double gmax = fmax(l12, fmax(l20, l01)) / nn;          // in-plane gradients of the three weights
    const double shift = S.delta_eff + S.theta * S.lam_max;
    const double E1p = E1 + 3.0 * gmax * shift;
    const double Kr = 48.0 * kPencilU * S.M + 3.0 * shift;
    double s = 0.0;
    for (int r = 0; r < 3; ++r) {
        double l1 = 0.0, d1 = 0.0;
        for (int k = 0; k < 3; ++k) {
            const double x = X0[r][k] + E0 * n[k];
            q[4 * r + k] = (float)x;
            l1 += fabs(x); d1 += fabs(dX[r][k]);
        }
        // 1.5*E1p*|n| + rounding: stored coefficients, dc = -(da + db) in float, the two FMAs per component that build X(E),
        // the three FMAs of the dot product and the constant itself -- <= 16u * (|X0'|_1 + 2*Emax*|dX|_1), |w| <= 1 + theta
        s = fmax(s, 1.5 * E1p * nn + 16.0 * kPencilU * (l1 + 2.0 * S.Emax * d1));
    }
    const float sf = pencil_round_up(s * 1.0001 + 1e-30);
    q[3] = sf;
    for (int k = 0; k < 3; ++k) { q[12 + k] = (float)da[k]; q[16 + k] = (float)db[k]; q[20 + k] = (float)n[k]; }
    const double c0 = v0[0] * n[0] + v0[1] * n[1] + v0[2] * n[2];
    q[15] = (float)c0;
    q[19] = pencil_round_up(nn * (1.0 + 4.0 * kPencilU));
    // distance clauses: K_r*|n| + rounding of det (<= 6u(|c0| + sqrt(3)*Emax*|n|), doubled) and of sigma_s*lam (8u*|n|*2*lam_max)
    q[23] = pencil_round_up((Kr * nn + 12.0 * kPencilU * (fabs(c0) + 1.7320508 * S.Emax * nn) + 16.0 * kPencilU * nn * 2.0 * S.lam_max) * 1.0001);
    q[7] = 0.0f; q[11] = 0.0f;   // id / nv: the caller's
    for (int i = 0; i < 24; ++i)
        if (!isfinite(q[i])) { tp_never(q); return false; }
    return true;
}

// ---- float side: identical arithmetic in the kernel and in the CPU replay ------------------------------------------------
#ifdef __CUDA_ARCH__
#define RT_TP_MUL(a, b) __fmul_rn((a), (b))
#define RT_TP_ADD(a, b) __fadd_rn((a), (b))
#define RT_TP_SUB(a, b) __fsub_rn((a), (b))
#define RT_TP_DIV(a, b) __fdiv_rn((a), (b))
#define RT_TP_SQRT(a) __fsqrt_rn((a))
#else
#define RT_TP_MUL(a, b) ((a) * (b))      // (host translation units that include this header are built with -ffp-contract=off)
#define RT_TP_ADD(a, b) ((a) + (b))
#define RT_TP_SUB(a, b) ((a) - (b))
#define RT_TP_DIV(a, b) ((a) / (b))
#define RT_TP_SQRT(a) sqrtf((a))
#endif

// Mirror image of the eye about the plane of triangle (A, B, C), centred coordinates.  Float; whatever its error, rays are
// checked against THIS point and the scan uses THIS point.  false: degenerate triangle.
RT_HD bool tp_mirror_point(const float eye[3], const float center[3], const float A[3], const float B[3], const float C[3], float E[3]) {
    const float ux = RT_TP_SUB(B[0], A[0]), uy = RT_TP_SUB(B[1], A[1]), uz = RT_TP_SUB(B[2], A[2]);
    const float vx = RT_TP_SUB(C[0], A[0]), vy = RT_TP_SUB(C[1], A[1]), vz = RT_TP_SUB(C[2], A[2]);
    const float nx = RT_TP_SUB(RT_TP_MUL(uy, vz), RT_TP_MUL(uz, vy)), ny = RT_TP_SUB(RT_TP_MUL(uz, vx), RT_TP_MUL(ux, vz)), nz = RT_TP_SUB(RT_TP_MUL(ux, vy), RT_TP_MUL(uy, vx));
    const float n2 = fmaf(nx, nx, fmaf(ny, ny, RT_TP_MUL(nz, nz)));
    if (!(n2 > 0.0f) || !(n2 < 1e30f)) return false;
    const float px = RT_TP_SUB(eye[0], A[0]), py = RT_TP_SUB(eye[1], A[1]), pz = RT_TP_SUB(eye[2], A[2]);
    const float t = RT_TP_DIV(RT_TP_MUL(2.0f, fmaf(px, nx, fmaf(py, ny, RT_TP_MUL(pz, nz)))), n2);
    E[0] = RT_TP_SUB(fmaf(-t, nx, eye[0]), center[0]);
    E[1] = RT_TP_SUB(fmaf(-t, ny, eye[1]), center[1]);
    E[2] = RT_TP_SUB(fmaf(-t, nz, eye[2]), center[2]);
    return true;
}

// Per-ray state of the filter.
struct TpRay { float wx, wy, wz, lam_o; };

// Acceptance check + ray set-up.  O, D: the continuation ray as built (un-centred); E: centred.  On success fills the unit direction
// and lam_o = |O - E|.
RT_HD bool tp_accepts(const TpSetup& S, const float E[3], const float O[3], const float D[3], TpRay& r) {
    const float dx = RT_TP_SUB(D[0], O[0]), dy = RT_TP_SUB(D[1], O[1]), dz = RT_TP_SUB(D[2], O[2]);
    const float ox = RT_TP_SUB(RT_TP_SUB(O[0], S.centerf[0]), E[0]), oy = RT_TP_SUB(RT_TP_SUB(O[1], S.centerf[1]), E[1]), oz = RT_TP_SUB(RT_TP_SUB(O[2], S.centerf[2]), E[2]);
    const float cx = RT_TP_SUB(RT_TP_MUL(oy, dz), RT_TP_MUL(oz, dy)), cy = RT_TP_SUB(RT_TP_MUL(oz, dx), RT_TP_MUL(ox, dz)), cz = RT_TP_SUB(RT_TP_MUL(ox, dy), RT_TP_MUL(oy, dx));
    const float c2 = fmaf(cx, cx, fmaf(cy, cy, RT_TP_MUL(cz, cz))), d2 = fmaf(dx, dx, fmaf(dy, dy, RT_TP_MUL(dz, dz))), o2 = fmaf(ox, ox, fmaf(oy, oy, RT_TP_MUL(oz, oz)));
    const float od = fmaf(ox, dx, fmaf(oy, dy, RT_TP_MUL(oz, dz)));
    const bool ok = (c2 <= RT_TP_MUL(S.half_delta2, d2)) && (od > 0.0f) && (o2 >= S.lam_min2) && (o2 < 1e30f) && (d2 > 1e-30f) && (d2 < 1e30f) &&
                    (fabsf(E[0]) <= S.emax) && (fabsf(E[1]) <= S.emax) && (fabsf(E[2]) <= S.emax);
    const float inv = RT_TP_DIV(1.0f, RT_TP_SQRT(d2));
    r.wx = RT_TP_MUL(dx, inv); r.wy = RT_TP_MUL(dy, inv); r.wz = RT_TP_MUL(dz, inv);
    r.lam_o = RT_TP_SQRT(o2);
    return ok;
}

// The thread's view of one triangle: weight vectors oriented by the side of the plane E lies on.
struct TpTri {
    float ax, ay, az, bx, by, bz, cx, cy, cz;   // oriented weight vectors
    float nx, ny, nz;                            // oriented normal (sigma_s = w.n_s > 0 for valid hits)
    float s, det_abs, krn, nlen;
    bool near_;                                  // the plane passes next to E: grazing pairs are possible
};
RT_HD float tp_flip(float x, uint32_t m) {
    uint32_t b;
    memcpy(&b, &x, 4);
    b ^= m;
    memcpy(&x, &b, 4);
    return x;
}
// q: the 24 floats of the record; E: centred common point; near_thr: (near_a * cos_g + near_b) of this thread.
RT_HD void tp_orient(const float q[24], const float E[3], float near_thr, TpTri& T) {
    const float dax = q[12], day = q[13], daz = q[14], dbx = q[16], dby = q[17], dbz = q[18];
    const float dcx = -RT_TP_ADD(dax, dbx), dcy = -RT_TP_ADD(day, dby), dcz = -RT_TP_ADD(daz, dbz);
    const float nEx = -E[0], nEy = -E[1], nEz = -E[2];
    // X(E) = X0' + E x dX,   E x d = (Ey dz - Ez dy, Ez dx - Ex dz, Ex dy - Ey dx)
    float ax = fmaf(E[1], daz, fmaf(nEz, day, q[0])), ay = fmaf(E[2], dax, fmaf(nEx, daz, q[1])), az = fmaf(E[0], day, fmaf(nEy, dax, q[2]));
    float bx = fmaf(E[1], dbz, fmaf(nEz, dby, q[4])), by = fmaf(E[2], dbx, fmaf(nEx, dbz, q[5])), bz = fmaf(E[0], dby, fmaf(nEy, dbx, q[6]));
    float cx = fmaf(E[1], dcz, fmaf(nEz, dcy, q[8])), cy = fmaf(E[2], dcx, fmaf(nEx, dcz, q[9])), cz = fmaf(E[0], dcy, fmaf(nEy, dcx, q[10]));
    const float det = fmaf(nEx, q[20], fmaf(nEy, q[21], fmaf(nEz, q[22], q[15])));   // (v0 - E).n
    uint32_t m;
    memcpy(&m, &det, 4);
    m &= 0x80000000u;
    T.ax = tp_flip(ax, m); T.ay = tp_flip(ay, m); T.az = tp_flip(az, m);
    T.bx = tp_flip(bx, m); T.by = tp_flip(by, m); T.bz = tp_flip(bz, m);
    T.cx = tp_flip(cx, m); T.cy = tp_flip(cy, m); T.cz = tp_flip(cz, m);
    T.nx = tp_flip(q[20], m); T.ny = tp_flip(q[21], m); T.nz = tp_flip(q[22], m);
    T.s = q[3]; T.det_abs = fabsf(det); T.krn = q[23]; T.nlen = q[19];
    T.near_ = T.det_abs < RT_TP_MUL(near_thr, T.nlen);   // (a "never" record has nlen = 0: never near; a NaN det: not near, weights NaN -> candidate)
}
// Hot test: sign word of the three weights (bit 31 set <=> certainly not a candidate, unless the triangle is near).
RT_HD uint32_t tp_weights(const TpTri& T, const TpRay& r) {
    const float a = fmaf(r.wx, T.ax, fmaf(r.wy, T.ay, fmaf(r.wz, T.az, T.s)));
    const float b = fmaf(r.wx, T.bx, fmaf(r.wy, T.by, fmaf(r.wz, T.bz, T.s)));
    const float c = fmaf(r.wx, T.cx, fmaf(r.wy, T.cy, fmaf(r.wz, T.cz, T.s)));
    uint32_t ua, ub, uc;
    memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); memcpy(&uc, &c, 4);
    return ua | ub | uc;
}
// Full test (cold path): weights + both distance clauses, or the grazing clause of a near triangle.  lam_hi = lam_o + nearest + slack
// (FLT_MAX: nothing yet), lam_lo = lam_o - slack, cgn = cos_g of the thread.
RT_HD bool tp_candidate(const TpTri& T, const TpRay& r, float lam_hi, float lam_lo, float cg) {
    const uint32_t w = tp_weights(T, r);
    const float sg = fmaf(r.wx, T.nx, fmaf(r.wy, T.ny, RT_TP_MUL(r.wz, T.nz)));
    const float e_hi = fmaf(sg, lam_hi, RT_TP_SUB(T.krn, T.det_abs));      // sigma_s*lam_hi - (|det| - KrN) >= 0
    const float e_lo = fmaf(-sg, lam_lo, RT_TP_ADD(T.det_abs, T.krn));     // (|det| + KrN) - sigma_s*lam_lo >= 0
    uint32_t uh, ul;
    memcpy(&uh, &e_hi, 4); memcpy(&ul, &e_lo, 4);
    const bool inside = !((w | uh | ul) >> 31);                            // NaN anywhere: sign bit clear -> candidate
    const bool grazing = T.near_ && !(fabsf(sg) > RT_TP_MUL(RT_TP_MUL(cg, 1.01f), T.nlen));
    return inside || grazing;
}
// cos_g and the near threshold of a thread whose rays have |O - E| >= lam_min_thread.
RT_HD void tp_thread_consts(const TpSetup& S, float lam_min_thread, float& cg, float& near_thr) {
    cg = fmaxf(S.cg_floor, RT_TP_DIV(S.cg_num, lam_min_thread));
    near_thr = fmaf(S.near_a, cg, S.near_b);
}

}  // namespace rt
