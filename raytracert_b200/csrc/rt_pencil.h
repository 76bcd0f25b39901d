// rt_pencil.h -- the "pencil" filter: rays whose lines all pass through one common point E.
//
// Primary rays all come from the eye (main.cpp:300-320 unprojects the same pixel on the near and the far plane, so the
// line origin -> dest passes through the eye up to rounding) and every shadow ray of a light ends AT the light
// (raytracing.cpp:246-248: dest = MyLightPositions[i], exactly).  For such a pencil of lines the ray-triangle test
// needs no ray origin at all.  With p_i = v_i - E and a ray direction w (pointing away from E):
//     a = w.(p1 x p2)   b = w.(p2 x p0)   c = w.(p0 x p1)        sigma = a + b + c = w.n,   n = (v1-v0) x (v2-v0)
// are the (unnormalised) barycentric weights of v0, v1, v2 at the point where the line meets the triangle's plane,
//     lambda = det / sigma,   det = p0.(p1 x p2)
// is that point's distance from E along w, and the line pierces the triangle iff a, b, c have the sign of sigma.
// Hits the reference can accept lie at lambda > 0 (in front of the eye / on the hit point's side of the light, see the
// launch conditions below), i.e. sign(sigma) = sign(det): the three vectors are pre-multiplied by sign(det), and the
// candidate test becomes "a', b', c' >= 0 and lambda < lambda_hi".
// Projective chart: all rays of a launch point into one half space (w.f >= 1/W_max for the launch's axis f: the view
// direction / the axis that separates the light from the scene box), the tests are homogeneous in w, so every ray is
// represented by w' = w / (w.f) = (x, y, 1) in an orthonormal frame (u, v, f) and the records are stored in that frame:
//     a = A'u*x + A'v*y + A'f        two FMAs per weight, the constant A'f + slack sits in the record
//     e = sigma * zeta_hi - det_lo,  sigma = Nu*x + Nv*y + Nf,  zeta = depth along f of the plane point (= lambda / |w'|)
// -- 9 packed FP32 instructions per (ray pair, triangle) instead of 16, no reciprocal, and the compare logic is the OR
// of four sign bits (2 LOP3 per ray).  The kernels' hot loop evaluates the three weights only (6 instructions); the
// distance clause e runs when the cold path rebuilds a block's candidate mask (rt_kernels.cuh: pencil_pair_hot / pencil_pair).
//
// This header is shared by the CUDA library (record construction in k_build_pencil, launch set-up in rt_b200.cu)
// and by the CPU soundness test (tests/pencil_check.cpp), which replays the filter with fmaf() -- the filter uses
// only IEEE FMAs, so the CPU replay is the same arithmetic as the FFMA2 instructions.
//
// Soundness ("pencil filter"): the tolerances E0, E1 bound the reference's own rounding relative to
// the true line through its float origin and dest (same constants as the generic filter record).  On top of that
//   * the true line misses E by at most `delta` and (x, y, 1) is within theta = 16u*w_max of its direction: the plane point moves by at most
//     (2.7/|cos|)(delta + theta*lam_max), a barycentric by gmax times that -> E1p = E1 + 3*gmax*(delta + theta*lam_max);
//   * a >= -(E0 + E1p/|cos|)*sigma  <=>  w'.(A + E0*n) + E1p*|n|*|w'| >= 0 : E0 is folded into the vector, E1p*|n|*w_max into
//     the constant term of the FMA chain together with the chain's own rounding (<= 4u(|A'||w'| + constant));
//   * distance: lambda < lam_O + best + s_lam + K_r/|cos|  <=>  |det| - K_r*|n| < zeta_hi * sigma with
//     zeta_hi = zeta_O + (best + s_lam)/|w'|, so the 1/|cos| part of the guard band is folded into the per-triangle
//     constant det_lo (sigma's own constant carries 8u|n|w_max: it is never under-estimated).
// Pairs with |cos| < cos_g need no answer: the pencil kernels are only used when the scene-level proof of
// rt_b200.cu:build_records says the reference rejects every such pair itself (no_grazing;
// cos_g = max(1.05e-5, 2.5*delta/lambda_min, 5*theta): E and the true line must be on the same side of the plane, and
// the chart direction must not change the cosine by more than a fifth).
// Exactness of the bound (no first-order argument): with H = signed distance of E to the plane, C = n.w, and H', C'
// the same for the true line, |H - H'| <= delta <= 0.4|H'| and |C - C'| <= theta <= 0.2|C'|, so lambda' = H/C has the
// sign of lambda* = H'/C', |lambda' - lambda*| <= (delta + lambda* theta)/(0.8|C'|), |X' - X*| <= 2.25(delta + lambda* theta)/|C'|,
// and 1/|C'| <= 1.2/|C| turns that into 2.7/|C| (the records use 3).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rt {

constexpr double kPencilU = 5.9604644775390625e-8;   // 2^-24

// ------------------------------------------------------------------------------------------------
// Tolerances of one triangle (shared by the generic filter record and the pencil record).
// W = dominant axis class of the record position (the generic filter's 2-D projection).
// ------------------------------------------------------------------------------------------------
struct FilterTol {
    bool always;            // the filter cannot bound this triangle ("always exact")
    double n3[3], nn;       // u x v and its length
    double su, sv, tu, tv;  // projected barycentric functionals (generic record)
    double E0, E1;
    int U, V;
};

RT_HD FilterTol filter_tolerances(const float A[3], const float B[3], const float C[3], int W, double M) {
    FilterTol t;
    t.always = false;
    const double u3[3] = {(double)B[0] - A[0], (double)B[1] - A[1], (double)B[2] - A[2]};
    const double v3d[3] = {(double)C[0] - A[0], (double)C[1] - A[1], (double)C[2] - A[2]};
    t.n3[0] = u3[1] * v3d[2] - u3[2] * v3d[1];
    t.n3[1] = u3[2] * v3d[0] - u3[0] * v3d[2];
    t.n3[2] = u3[0] * v3d[1] - u3[1] * v3d[0];
    t.nn = sqrt(t.n3[0] * t.n3[0] + t.n3[1] * t.n3[1] + t.n3[2] * t.n3[2]);
    const double uu = u3[0] * u3[0] + u3[1] * u3[1] + u3[2] * u3[2], vv = v3d[0] * v3d[0] + v3d[1] * v3d[1] + v3d[2] * v3d[2];
    t.U = (W + 1) % 3; t.V = (W + 2) % 3;
    const double det = u3[t.U] * v3d[t.V] - u3[t.V] * v3d[t.U];   // == n3[W]
    t.su = t.sv = t.tu = t.tv = 0.0; t.E0 = t.E1 = 0.0;
    if (!(t.nn > 0.0) || !isfinite(t.nn) || !(uu > 0.0) || !(vv > 0.0) || !(fabs(det) > 0.0) || !isfinite(det)) { t.always = true; return t; }
    // s = 1 at B, t = 1 at C, both 0 at A, as functions of the (U, V) coordinates
    t.su = v3d[t.V] / det; t.sv = -v3d[t.U] / det; t.tu = -u3[t.V] / det; t.tv = u3[t.U] / det;
    const double gs = sqrt(t.su * t.su + t.sv * t.sv), gt = sqrt(t.tu * t.tu + t.tv * t.tv);
    const double gq = sqrt((t.su + t.tu) * (t.su + t.tu) + (t.sv + t.tv) * (t.sv + t.tv));
    const double gmax = fmax(gs, fmax(gt, gq));   // >= the in-plane gradients (the projection only stretches them)
    const double sinphi = t.nn / sqrt(uu * vv);
    const double kappa = fmax(1.0, 0.25 / sinphi);
    // DESIGN.md "filter soundness": E0 covers the rounding of the reference's own dot-product barycentrics
    // (<= 28uM*gmax/sin(phi) + 8u/sin^2(phi)) plus the generic filter's arithmetic (<= 16uM*gmax),
    // E1*|1/cos| the shift of the plane point caused by the two sides' error along the ray (<= 34uM/|cos|)
    t.E0 = 256.0 * kPencilU * M * gmax * kappa + 1e-6;
    t.E1 = 64.0 * kPencilU * M * gmax * kappa;
    if (!(t.E0 < 64.0) || !isfinite(t.E0)) t.always = true;   // beyond this the dilated triangle is so large that "always exact" is cheaper
    return t;
}

// ------------------------------------------------------------------------------------------------
// Pencil launch set-up (one per common point: the camera, or one light)
// ------------------------------------------------------------------------------------------------
struct PencilSetup {
    double E[3];      // the common point
    double M;         // power-of-two bound on |coordinate| of the scene, the ray origins and E
    double delta;     // every ray's true line (through its float origin and dest) passes within delta of E
    double lam_max;   // >= |X - E| for every scene point X
    double cos_g;     // the filter answers for pairs with |cos(true line, plane)| >= cos_g; below that the launch must
                      // be covered by the scene-level proof that the reference rejects the pair itself
    double fu[3], fv[3], ff[3];   // orthonormal chart frame; every ray of the launch has w.ff >= 1/w_max
    double w_max;     // >= |w'| = 1/(w.ff) over the launch's rays (rays beyond it are handled exactly by the kernel)
    double theta;     // angular error of the float chart coordinates (x, y, 1) against the true direction
    float lam_slack;  // s_lam: ray-side guard band of the distance test
    float Ef[3];      // E rounded to float (what the kernels subtract)
    float F[9];       // frame rounded to float: u, v, f
    float w_max2;     // w_max^2 (1 + x^2 + y^2 must not exceed it)
};
constexpr double kPencilCosMin = 1.05e-5;   // the generic filter's grazing threshold (cos_min = 1e-5) + its 8u evaluation error

// M_scene: bound on |coordinate| of the scene and the ray origins.  Sets M (also covers E), lam_max and s_lam.
RT_HD void pencil_finish_setup(PencilSetup& S, double M_scene) {
    double Me = M_scene, e2 = 0.0;
    for (int k = 0; k < 3; ++k) { Me = fmax(Me, fabs(S.E[k])); e2 += S.E[k] * S.E[k]; S.Ef[k] = (float)S.E[k]; }
    S.M = Me;
    if (!(S.lam_max > 0.0)) S.lam_max = 1.7320508 * M_scene + sqrt(e2) + 1e-3 * Me;   // |X - E| <= |X| + |E| (the callers may know better)
    // s_lam = 128uM (reference side + evaluation of the origin's depth) + 3*delta, rounded up
    S.lam_slack = (float)((128.0 * kPencilU * S.M + 3.0 * S.delta) * 1.0001);
    for (int k = 0; k < 3; ++k) { S.F[k] = (float)S.fu[k]; S.F[3 + k] = (float)S.fv[k]; S.F[6 + k] = (float)S.ff[k]; }
    // x = (dir.u)/(dir.f) in float: |dx| <= 4u(1 + |x|)|w'| + u|x| (3-term dots of a rounded difference with a rounded
    // frame vector, one division), so the direction (x, y, 1) is within 16u*w_max of the true one
    S.theta = 16.0 * kPencilU * S.w_max;
    S.w_max2 = (float)(S.w_max * S.w_max * 1.00001);
}

// Completes an orthonormal frame around the unit axis f.
RT_HD void pencil_frame(PencilSetup& S, const double f[3]) {
    int k0 = 0;
    if (fabs(f[1]) < fabs(f[k0])) k0 = 1;
    if (fabs(f[2]) < fabs(f[k0])) k0 = 2;
    double t[3] = {0.0, 0.0, 0.0};
    t[k0] = 1.0;
    double u[3] = {f[1] * t[2] - f[2] * t[1], f[2] * t[0] - f[0] * t[2], f[0] * t[1] - f[1] * t[0]};
    const double ul = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (int k = 0; k < 3; ++k) { u[k] /= ul; S.ff[k] = f[k]; S.fu[k] = u[k]; }
    S.fv[0] = f[1] * u[2] - f[2] * u[1]; S.fv[1] = f[2] * u[0] - f[0] * u[2]; S.fv[2] = f[0] * u[1] - f[1] * u[0];
}

// lam_max from the box of everything a record can make a hit of (+ the 0.1 shadow bias and slack): farthest corner from E.
RT_HD double pencil_farthest_corner(const double E[3], const float lo[3], const float hi[3]) {
    double m2 = 0.0;
    for (int k = 0; k < 3; ++k) {
        const double a = fabs((double)lo[k] - 0.2 - E[k]), b = fabs((double)hi[k] + 0.2 - E[k]);
        const double m = fmax(a, b);
        m2 += m * m;
    }
    return sqrt(m2);
}

// Pencil record (4 float4 = 64 B, same position / tile layout as the generic record), vectors in the chart frame:
//   q0 = ( A'u, A'v, A'f + sA, Nu )        a'' = fma(A'u, x, fma(A'v, y, A'f + sA))      weight of v0 (1 - s - t)
//   q1 = ( B'u, B'v, B'f + sB, Nv )                                                        weight of v1 (s)
//   q2 = ( C'u, C'v, C'f + sC, Nf + sN )   sigma = fma(Nu, x, fma(Nv, y, Nf + sN))        N = sign(det) * n in the chart frame
//   q3 = ( -det_lo, id, nv, 0 )            e = fma(sigma, zeta_hi, -det_lo);  id / nv copied from the generic record
// candidate  <=>  sign bits of a'', b'', c'', e all clear.   "never" record: vectors 0, constants -1.
// (sigma gets its own two FMAs instead of a'' + b'' + c'': the weights carry the whole tolerance in their constants,
// and that sum would loosen the distance clause by the same relative amount.)
RT_HD void pencil_never(float q[16]) {
    for (int i = 0; i < 12; ++i) q[i] = 0.0f;
    q[2] = q[6] = q[10] = q[11] = -1.0f;
    q[12] = 0.0f; q[15] = 0.0f;   // q[13] (id), q[14] (nv) are the caller's
}

RT_HD float pencil_round_up(double v) {   // a float >= v
    float f = (float)v;
    if ((double)f < v) f = (f > 0.0f) ? f * 1.0000002f + 1e-37f : f * 0.9999998f + 1e-37f;
    return f;
}

// Launches WITHOUT the scene-level clause-free proof (RT_OPT_PENCIL_ANY, experimental): a pair that hits inside the scene
// has |cos| = |H'|/lambda >= |H'|/lam_max (H' = distance of the true line's point nearest E to the plane), so for a
// triangle whose plane stays lam_max*cos_g + 2*delta away from E no pair below cos_g exists, by geometry alone.  The (few)
// triangles nearer than that get an "always candidate" record: the exact path decides for every ray.
RT_HD bool pencil_plane_near(const float A[3], const float B[3], const float C[3], const PencilSetup& S) {
    const double u[3] = {(double)B[0] - A[0], (double)B[1] - A[1], (double)B[2] - A[2]};
    const double v[3] = {(double)C[0] - A[0], (double)C[1] - A[1], (double)C[2] - A[2]};
    const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double H = ((double)A[0] - S.E[0]) * n[0] + ((double)A[1] - S.E[1]) * n[1] + ((double)A[2] - S.E[2]) * n[2];
    return !(fabs(H) >= (S.lam_max * S.cos_g + 2.0 * S.delta) * nn * 1.001);   // NaN: near
}
RT_HD void pencil_always(float q[16]) {   // weights = 1, sigma = 1, e = zeta_hi + 1e30 > 0: a candidate for every ray
    for (int i = 0; i < 12; ++i) q[i] = 0.0f;
    q[2] = q[6] = q[10] = q[11] = 1.0f;
    q[12] = 1e30f; q[15] = 0.0f;
}
constexpr unsigned int kPencilMaxNear = 16;   // more "always candidate" records than this: the launch keeps the generic kernels

// Returns false (and writes a "never" record) when no ray of the pencil can validly hit the triangle.
RT_HD bool pencil_record(const float A[3], const float B[3], const float C[3], double E0, double E1, const PencilSetup& S, float q[16]) {
    double p0[3], p1[3], p2[3], u[3], v[3], e12[3];
    for (int k = 0; k < 3; ++k) {
        p0[k] = (double)A[k] - S.E[k]; p1[k] = (double)B[k] - S.E[k]; p2[k] = (double)C[k] - S.E[k];
        u[k] = (double)B[k] - A[k]; v[k] = (double)C[k] - A[k]; e12[k] = (double)C[k] - B[k];
    }
    const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double det = p0[0] * n[0] + p0[1] * n[1] + p0[2] * n[2];
    const double lp0 = sqrt(p0[0] * p0[0] + p0[1] * p0[1] + p0[2] * p0[2]);
    // E (numerically) in the triangle's plane: a valid non-grazing hit would need lambda*|cos| = |det|/|n| <= 1e-9|p0|,
    // i.e. lambda <= 1e-4*|p0| -- nearer to E than any hit the launch conditions allow (lambda_min >= 1e-3*M).
    // Also the sign of det (computed to ~1e-15 relative to |p0||n|) is reliable beyond this threshold.
    if (!(nn > 0.0) || !isfinite(nn) || !isfinite(det) || !(fabs(det) > 1e-9 * lp0 * nn)) { pencil_never(q); return false; }
    const double sg = det > 0.0 ? 1.0 : -1.0;
    double Av[3] = {p1[1] * p2[2] - p1[2] * p2[1], p1[2] * p2[0] - p1[0] * p2[2], p1[0] * p2[1] - p1[1] * p2[0]};
    double Bv[3] = {p2[1] * p0[2] - p2[2] * p0[1], p2[2] * p0[0] - p2[0] * p0[2], p2[0] * p0[1] - p2[1] * p0[0]};
    double Cv[3] = {p0[1] * p1[2] - p0[2] * p1[1], p0[2] * p1[0] - p0[0] * p1[2], p0[0] * p1[1] - p0[1] * p1[0]};
    // in-plane gradients of the three weights: |opposite edge| / |n|
    const double l12 = sqrt(e12[0] * e12[0] + e12[1] * e12[1] + e12[2] * e12[2]);
    const double lu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), lv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double gmax = fmax(l12, fmax(lu, lv)) / nn;
    const double shift = S.delta + S.theta * S.lam_max;                 // how far the pencil line can be from the true line near the scene
    const double E1p = E1 + 3.0 * gmax * shift;
    const double Kr = 48.0 * kPencilU * S.M + 3.0 * shift;              // 1/|cos| part of the distance guard band
    double* vec[3] = {Av, Bv, Cv};
    for (int r = 0; r < 3; ++r) {
        double x[3], len2 = 0.0;
        for (int k = 0; k < 3; ++k) { x[k] = sg * (vec[r][k] + E0 * n[k]); len2 += x[k] * x[k]; }
        const double cu = x[0] * S.fu[0] + x[1] * S.fu[1] + x[2] * S.fu[2];
        const double cv = x[0] * S.fv[0] + x[1] * S.fv[1] + x[2] * S.fv[2];
        const double cf = x[0] * S.ff[0] + x[1] * S.ff[1] + x[2] * S.ff[2];
        // constant term: E1p*|n|*|w'| (x1.5) + rounding of the two FMAs, of the stored coefficients and of the constant
        // itself (<= 4u(|A'||w'| + sA), x2), |w'| <= w_max
        const double sA = (1.5 * E1p * nn + 8.0 * kPencilU * sqrt(len2)) * S.w_max * 1.0001 + 1e-30;
        q[4 * r] = (float)cu; q[4 * r + 1] = (float)cv;
        q[4 * r + 2] = pencil_round_up(cf + sA);
        if (!isfinite(q[4 * r]) || !isfinite(q[4 * r + 1]) || !isfinite(q[4 * r + 2])) { pencil_never(q); return false; }
    }
    {   // sigma = w'.N must not come out below the true value: constant + 4u|n||w'| (two FMAs, stored coefficients, the constant)
        const double nu = sg * (n[0] * S.fu[0] + n[1] * S.fu[1] + n[2] * S.fu[2]);
        const double nv = sg * (n[0] * S.fv[0] + n[1] * S.fv[1] + n[2] * S.fv[2]);
        const double nf = sg * (n[0] * S.ff[0] + n[1] * S.ff[1] + n[2] * S.ff[2]);
        q[3] = (float)nu; q[7] = (float)nv;
        q[11] = pencil_round_up(nf + 8.0 * kPencilU * nn * S.w_max + 1e-30);
        if (!isfinite(q[3]) || !isfinite(q[7]) || !isfinite(q[11])) { pencil_never(q); return false; }
    }
    // det_lo <= (|det| - Kr*|n|) * (1 - 8u), rounded down; negative is fine (the distance clause then always passes)
    const double dl = fabs(det) - Kr * nn;
    double dlo = dl > 0.0 ? dl * (1.0 - 16.0 * kPencilU) : dl * (1.0 + 16.0 * kPencilU);
    float f = (float)dlo;
    if ((double)f > dlo) f = f > 0.0f ? f * 0.9999998f : f * 1.0000002f - 1e-37f;
    q[12] = -f;
    q[15] = 0.0f;
    return true;
}

// ------------------------------------------------------------------------------------------------
// Host side: can the primary rays of a frame be treated as a pencil, and around which point?
// corners: the 24 floats of rt_params (4 x (origin, dest): c00, c01, c10, c11 -- main.cpp:355-358).
// Every primary ray is the bilinear blend (weights w_i >= 0, sum 1) of the four corner rays, origin and dest with the
// SAME weights (main.cpp:380-386).  With o_i = O_i - E, d_i = D_i - E:
//     (O - E) x (D - E) = sum_i w_i^2 (o_i x d_i) + sum_{i<j} w_i w_j (o_i x d_j + o_j x d_i)
// is bounded by max(|o_i x d_i|, |o_i x d_j + o_j x d_i| / 2) because sum_i w_i^2 + 2 sum_{i<j} w_i w_j = 1, and
// |D - O| >= min_i (d_i - o_i).g for the unit mean direction g.  Their ratio bounds the distance from E to the exact
// blended line; the float evaluation of the blend moves origin and dest by <= eps_o each.
// Returns false when the frame is not a (forward) pencil: parallel rays, eye behind the near plane, ...
// ------------------------------------------------------------------------------------------------
inline bool pencil_camera_setup(const float corners[24], double M_scene, const float* box_lo, const float* box_hi, PencilSetup& S) {
    double O[4][3], D[4][3], g[4][3];
    double Mc = 0.0;
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 3; ++k) {
            O[c][k] = corners[6 * c + k]; D[c][k] = corners[6 * c + 3 + k]; g[c][k] = D[c][k] - O[c][k];
            if (!isfinite(O[c][k]) || !isfinite(D[c][k])) return false;
            Mc = fmax(Mc, fmax(fabs(O[c][k]), fabs(D[c][k])));
        }
    // least-squares point closest to the four corner lines: sum_i (I - g_i g_i^T) E = sum_i (I - g_i g_i^T) O_i
    double Am[3][3] = {{0}}, bv[3] = {0, 0, 0};
    for (int c = 0; c < 4; ++c) {
        const double l2 = g[c][0] * g[c][0] + g[c][1] * g[c][1] + g[c][2] * g[c][2];
        if (!(l2 > 0.0)) return false;
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) {
                const double m = (r == k ? 1.0 : 0.0) - g[c][r] * g[c][k] / l2;
                Am[r][k] += m; bv[r] += m * O[c][k];
            }
    }
    const double det = Am[0][0] * (Am[1][1] * Am[2][2] - Am[1][2] * Am[2][1]) - Am[0][1] * (Am[1][0] * Am[2][2] - Am[1][2] * Am[2][0]) +
                       Am[0][2] * (Am[1][0] * Am[2][1] - Am[1][1] * Am[2][0]);
    if (!(fabs(det) > 1e-9)) return false;   // (near-)parallel rays: no common point
    double E[3];
    for (int k = 0; k < 3; ++k) {
        double Mk[3][3];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Mk[r][c] = (c == k) ? bv[r] : Am[r][c];
        E[k] = (Mk[0][0] * (Mk[1][1] * Mk[2][2] - Mk[1][2] * Mk[2][1]) - Mk[0][1] * (Mk[1][0] * Mk[2][2] - Mk[1][2] * Mk[2][0]) +
                Mk[0][2] * (Mk[1][0] * Mk[2][1] - Mk[1][1] * Mk[2][0])) / det;
        if (!isfinite(E[k])) return false;
    }
    double o[4][3], d[4][3];
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 3; ++k) { o[c][k] = O[c][k] - E[k]; d[c][k] = D[c][k] - E[k]; }
    auto cross = [](const double* a, const double* b, double* r) { r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0]; };
    auto norm = [](const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
    double crossmax = 0.0;
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) {
            double x[3], y[3];
            cross(o[i], d[j], x);
            if (i == j) { crossmax = fmax(crossmax, norm(x)); continue; }
            cross(o[j], d[i], y);
            const double s[3] = {x[0] + y[0], x[1] + y[1], x[2] + y[2]};
            crossmax = fmax(crossmax, 0.5 * norm(s));
        }
    double gm[3] = {0, 0, 0};
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 3; ++k) gm[k] += g[c][k];
    const double gl = norm(gm);
    if (!(gl > 0.0)) return false;
    for (int k = 0; k < 3; ++k) gm[k] /= gl;
    double lenmin = INFINITY, lam_o_min = INFINITY;
    for (int c = 0; c < 4; ++c) lenmin = fmin(lenmin, g[c][0] * gm[0] + g[c][1] * gm[1] + g[c][2] * gm[2]);
    if (!(lenmin > 0.0)) return false;
    // forward pencil: both O - E = sum w_i o_i and D - O = sum w_i g_i have a positive component along the mean direction
    // (o_i.gm > 0, g_i.gm > 0 for every corner) and are collinear up to delta, so they point the same way; and
    // lambda_O = |O - E| >= (O - E).gm >= min_i o_i.gm
    for (int c = 0; c < 4; ++c) lam_o_min = fmin(lam_o_min, o[c][0] * gm[0] + o[c][1] * gm[1] + o[c][2] * gm[2]);
    if (!(M_scene < 1e18)) return false;
    double Me = M_scene, omax = 0.0;
    for (int k = 0; k < 3; ++k) { Me = fmax(Me, fabs(E[k])); S.E[k] = E[k]; }
    for (int c = 0; c < 4; ++c) omax = fmax(omax, norm(o[c]));
    // float evaluation of the blend: <= 4u*Mc per component on origin and dest, plus the common scaling (1 + eta),
    // |eta| <= 2u, of both about the world origin (the float weights do not sum to exactly 1).  Moving origin and dest
    // by eps moves the line's point at parameter tau (0 at the origin, 1 at dest) by <= (|1 - tau| + |tau|)*eps; E sits
    // at tau = -lambda_O / |D - O|, |tau| <= omax / lenmin.
    const double eps_o = 4.0 * kPencilU * 1.7320508 * Mc + 2.0 * kPencilU * 1.7320508 * fmax(Mc, Me);
    S.delta = crossmax / lenmin + (1.0 + 2.0 * omax / lenmin) * eps_o;
    // chart: f = mean view direction.  (x, y) of a blended ray is a convex combination of the corner rays' (x_i, y_i)
    // (weights w_i g_i.f > 0), so |w'| <= max over the corners
    pencil_frame(S, gm);
    S.w_max = 0.0;
    for (int c = 0; c < 4; ++c) {
        const double gf = g[c][0] * gm[0] + g[c][1] * gm[1] + g[c][2] * gm[2];
        if (!(gf > 0.0)) return false;
        S.w_max = fmax(S.w_max, norm(g[c]) / gf);
    }
    S.w_max *= 1.0 + 1e-5;
    if (!(S.w_max < 8.0)) return false;      // wider than ~83 degrees off axis: not worth it
    S.lam_max = (box_lo && box_hi && box_lo[0] <= box_hi[0] && box_lo[1] <= box_hi[1] && box_lo[2] <= box_hi[2]) ? pencil_farthest_corner(E, box_lo, box_hi) : 0.0;
    pencil_finish_setup(S, M_scene);
    // launch conditions: E well in front of every ray origin (lambda_min >= 2e-3*M, see pencil_record) and a usable delta
    const double lam_min = lam_o_min - 4.0 * S.delta - 64.0 * kPencilU * S.M;
    if (!(lam_min >= 2e-3 * S.M)) return false;
    if (!(S.delta <= 1e-3 * S.M)) return false;
    // E and the true line's closest point E' must lie on the same side of every plane a valid pair can hit:
    // |dist(E', plane)| = lambda*|cos| >= lam_min*cos_g must exceed delta with room to spare (factor 2.5)
    // and the chart direction must keep the cosine's sign and size: theta <= 0.2*cos_g
    S.cos_g = fmax(kPencilCosMin, fmax(2.5 * S.delta / lam_min, 5.0 * S.theta));
    return true;
}

// Shadow rays of one light: dest = the light exactly, so delta = 0.  box_lo / box_hi: the bounding box of everything a
// filter record can make a hit of (union of the tile boxes); box' = box + 0.2 also holds the ray origins (hit + 0.1
// bias) of every hit inside the box.  Chart axis f = direction from the light to the centre of the box; the launch
// needs the whole box' strictly inside the half space (X - L).f > 0.  Directions run from the light to the ray origin.
// The pencil handles occluders on the origin's side of the light (0 < depth <= depth of the origin).  A ray whose
// origin is not in the half space (its continuation BEYOND the light could re-enter the box: the reference's shadow
// rays are unbounded) or lies outside the chart (|w'| > w_max) is handled exactly by the kernel (k_shadow, "unsafe"
// rays); origins inside box' never are: X -> (X - L)/((X - L).f) maps box' to a convex polygon of the chart plane, so
// |w'| is largest at a corner.
inline bool pencil_light_setup(const float L[3], const float box_lo[3], const float box_hi[3], double M_scene, PencilSetup& S) {
    double Me = M_scene;
    for (int k = 0; k < 3; ++k) {
        if (!isfinite(L[k]) || !(box_lo[k] <= box_hi[k])) return false;   // empty / NaN box
        Me = fmax(Me, fabs((double)L[k]));
    }
    if (!(Me < 1e18)) return false;
    double f[3], fl = 0.0;
    for (int k = 0; k < 3; ++k) { f[k] = 0.5 * ((double)box_lo[k] + box_hi[k]) - L[k]; fl += f[k] * f[k]; }
    fl = sqrt(fl);
    if (!(fl > 0.0) || !isfinite(fl)) return false;
    for (int k = 0; k < 3; ++k) f[k] /= fl;
    double gap = INFINITY, wmax = 0.0;
    for (int c = 0; c < 8; ++c) {
        double X[3], l2 = 0.0, dp = 0.0;
        for (int k = 0; k < 3; ++k) { X[k] = (((c >> k) & 1) ? (double)box_hi[k] + 0.2 : (double)box_lo[k] - 0.2) - L[k]; l2 += X[k] * X[k]; dp += X[k] * f[k]; }
        gap = fmin(gap, dp);
        if (dp > 0.0) wmax = fmax(wmax, sqrt(l2) / dp);
    }
    if (!(gap >= 2e-3 * Me)) return false;            // the light is inside (or too close to) the box: no pencil
    for (int k = 0; k < 3; ++k) S.E[k] = L[k];
    S.delta = 0.0;
    S.lam_max = pencil_farthest_corner(S.E, box_lo, box_hi);
    S.w_max = wmax * (1.0 + 1e-5);
    if (!(S.w_max < 8.0)) return false;               // too wide a cone for a useful chart
    pencil_frame(S, f);
    pencil_finish_setup(S, M_scene);
    S.cos_g = fmax(kPencilCosMin, 5.0 * S.theta);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Reflection pencils: the continuation rays of PRIMARY hits on one plane n.X = d (|n| = 1) -- a floor, a wall, the
// water of the stand-in scene: many coplanar triangles -- leave the mirror image E* of the eye, because reflection()
// (raytracing.cpp:277-285) mirrors a line through the eye about that plane.  The chart is the mirrored camera chart.
// Nothing here relies on an error analysis of the reflection arithmetic: k_shade routes a continuation ray to this
// pencil only after pencil_mirror_accepts() has CHECKED, in float, that its line passes within delta/2 of E*, that
// its direction lies in the chart and that E* is at least lam_min behind its origin; every other ray takes the generic
// scan.  delta >= 4*delta_cam + 64u*max(M, lam_max) keeps the check's own rounding (<= 32u*lam_max) inside the other half.
// ------------------------------------------------------------------------------------------------
struct MirrorCheck {       // what pencil_mirror_accepts needs, 16 floats
    float E[3], half_delta2;   // E* (float), (delta / 2)^2
    float F[9];                // chart frame u, v, f (float)
    float w_max2, lam_min2, pad;
};

inline bool pencil_mirror_setup(const PencilSetup& cam, const double n[3], double d, double M_scene, const float* box_lo, const float* box_hi,
                                PencilSetup& S, MirrorCheck& C) {
    const double nl = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (!(fabs(nl - 1.0) < 1e-6) || !isfinite(d)) return false;
    const double sd = n[0] * cam.E[0] + n[1] * cam.E[1] + n[2] * cam.E[2] - d;   // signed distance of the eye to the plane
    const double h = fabs(sd);
    auto mirror_vec = [&](const double v[3], double out[3]) {
        const double t = 2.0 * (v[0] * n[0] + v[1] * n[1] + v[2] * n[2]);
        for (int k = 0; k < 3; ++k) out[k] = v[k] - t * n[k];
    };
    for (int k = 0; k < 3; ++k) S.E[k] = cam.E[k] - 2.0 * sd * n[k];
    double f[3];
    mirror_vec(cam.ff, f);
    const double fl = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int k = 0; k < 3; ++k) f[k] /= fl;
    pencil_frame(S, f);      // any orthonormal completion will do: the chart bound is rotation invariant
    S.w_max = cam.w_max * (1.0 + 1e-4);
    if (!(S.w_max < 8.0) || !(M_scene < 1e18)) return false;
    S.lam_max = (box_lo && box_hi && box_lo[0] <= box_hi[0] && box_lo[1] <= box_hi[1] && box_lo[2] <= box_hi[2]) ? pencil_farthest_corner(S.E, box_lo, box_hi) : 0.0;
    S.delta = 0.0;
    pencil_finish_setup(S, M_scene);   // S.M, lam_max (if 0), theta, frame floats
    S.delta = 4.0 * cam.delta + 64.0 * kPencilU * fmax(S.M, S.lam_max);
    S.lam_slack = (float)((128.0 * kPencilU * S.M + 3.0 * S.delta) * 1.0001);
    // the origins sit on the plane (+ the 0.01 offset along the reflected ray): at least ~h from E*; half of it is demanded
    const double lam_min = 0.5 * h;
    if (!(lam_min >= 2e-3 * S.M) || !(S.delta <= 1e-3 * S.M)) return false;
    S.cos_g = fmax(kPencilCosMin, fmax(2.5 * S.delta / lam_min, 5.0 * S.theta));
    for (int k = 0; k < 3; ++k) C.E[k] = S.Ef[k];
    for (int k = 0; k < 9; ++k) C.F[k] = S.F[k];
    C.half_delta2 = (float)(0.25 * S.delta * S.delta * 0.999);
    C.w_max2 = S.w_max2;
    C.lam_min2 = (float)(lam_min * lam_min * 1.001);
    C.pad = 0.f;
    return true;
}

// Chart coordinates of a ray in the frame F (pencil_set_slot, k_shade and the CPU replay share this): false = not representable.
RT_HD bool pencil_chart_xy(const float F[9], float w_max2, float dx, float dy, float dz, float& x, float& y) {
    const float den = fmaf(dx, F[6], fmaf(dy, F[7], dz * F[8]));
#ifdef __CUDA_ARCH__
    const float inv = __fdiv_rn(1.0f, den);
#else
    const float inv = 1.0f / den;
#endif
    x = fmaf(dx, F[0], fmaf(dy, F[1], dz * F[2])) * inv;
    y = fmaf(dx, F[3], fmaf(dy, F[4], dz * F[5])) * inv;
    return (den > 0.0f) && (fmaf(x, x, fmaf(y, y, 1.0f)) <= w_max2);     // NaN / inf fail both
}

// Does the line through O and D (the continuation ray k_shade just built) belong to the mirror pencil?  Float arithmetic;
// |(O - E*) x dir| is evaluated to within 8u|O - E*||dir|, i.e. the distance to within 32u*lam_max < delta/2.
RT_HD bool pencil_mirror_accepts(const MirrorCheck& C, const float O[3], const float D[3]) {
    const float dx = D[0] - O[0], dy = D[1] - O[1], dz = D[2] - O[2];
    const float ox = O[0] - C.E[0], oy = O[1] - C.E[1], oz = O[2] - C.E[2];
    float x, y;
    if (!pencil_chart_xy(C.F, C.w_max2, dx, dy, dz, x, y)) return false;
    const float cx = oy * dz - oz * dy, cy = oz * dx - ox * dz, cz = ox * dy - oy * dx;
    const float c2 = cx * cx + cy * cy + cz * cz, d2 = dx * dx + dy * dy + dz * dz, o2 = ox * ox + oy * oy + oz * oz;
    const float od = ox * dx + oy * dy + oz * dz;
    return (c2 <= C.half_delta2 * d2) && (od > 0.0f) && (o2 >= C.lam_min2) && (o2 < 1e30f);
}

}  // namespace rt
