// pencil_check.cpp -- TEST INFRASTRUCTURE: CPU replay of the pencil filter (raytracert_b200/csrc/rt_pencil.h) and a CPU
// restatement of the generic filter (generic_check, at the end).
//
// The pencil filter uses only IEEE FMAs and adds, so fmaf() here is the arithmetic of the FFMA2 / FADD2 instructions
// of k_trace / k_shadow.  For a batch of rays and a triangle soup this replays record construction (the same
// pencil_record() the CUDA library compiles) and the candidate test for EVERY (ray, triangle) pair and compares with
// the oracle's decision for the pair (a function pointer to oracle/rt_oracle.c:orc_ray_triangle handed in by the test):
// a pair the reference accepts must be a candidate -- for the worst admissible "nearest so far" (the next float above
// the pair's own distance) in nearest-hit mode, and for r >= 0 in any-hit (shadow) mode.
//
// Build (tests/test_pencil_filter.py): g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../raytracert_b200/csrc/rt_pencil.h"
#include "../raytracert_b200/csrc/rt_tpencil.h"

using namespace rt;

typedef int (*pair_fn_t)(const float*, const float*, const float*, const float*, const float*, float*);

struct PencilCheckResult {
    int64_t pairs, ref_hits, candidates, violations, grazing_skipped, unsafe_rays, always_tris, never_recs;
    int32_t setup_ok, first_bad_ray, first_bad_tri, pad;
    double delta, M, cos_g;
};

static float round_up_sum(float a, float b) {   // __fadd_ru
    const double s = (double)a + (double)b;
    float f = (float)s;
    if ((double)f < s) f = nextafterf(f, INFINITY);
    return f;
}

static int dominant_axis(const float* A, const float* B, const float* C) {   // rt_upload_scene's class
    const double u[3] = {(double)B[0] - A[0], (double)B[1] - A[1], (double)B[2] - A[2]};
    const double v[3] = {(double)C[0] - A[0], (double)C[1] - A[1], (double)C[2] - A[2]};
    const double nx = std::fabs(u[1] * v[2] - u[2] * v[1]), ny = std::fabs(u[2] * v[0] - u[0] * v[2]), nz = std::fabs(u[0] * v[1] - u[1] * v[0]);
    int w = 0;
    if (ny > nx) w = 1;
    if (nz > (w == 1 ? ny : nx)) w = 2;
    return w;
}

// g_premise = 1: the launch relies on the scene-level clause-free proof (pairs with |cos| < cos_g are certain misses in the
// reference; what rt_b200.cu requires today).  g_premise = 0: no such proof -- instead a triangle whose plane passes within
// lam_max*cos_g + 2*delta of the common point gets an "always candidate" record (counted in near_plane), because for
// every other triangle a pair that can hit inside the scene has |cos| = |H'|/lambda >= cos_g by geometry alone.
static int g_premise = 1;
static long long g_near_plane = 0;

extern "C" {

void pencil_check_set_premise(int on) { g_premise = on; }
static double g_plane[4] = {0, 1, 0, 0};
void pencil_check_set_plane(double nx, double ny, double nz, double d) { g_plane[0] = nx; g_plane[1] = ny; g_plane[2] = nz; g_plane[3] = d; }
long long pencil_check_near_planes(void) { return g_near_plane; }

// mode 2 (groundwork for reflection pencils): rays leaving a common point that the caller describes
// directly -- setup24 = E (3 floats), chart axis f (3, any length), delta, w_max, lam_min; nearest-hit semantics like mode 0.
// mode 0: primary rays of the camera `setup24` (24 corner floats); mode 1: shadow rays ending at the light `setup24[0..2]`,
// box = setup24[3..5] (lo) / [6..8] (hi).  tri: ntri x 9 floats.  rays: n x 6 floats (origin, dest).
// inv_scale perturbs the chart division (x, y scaled by it: a few ulp of extra direction error).
int pencil_check(int mode, const float* setup24, double M_scene, int ntri, const float* tri, int nrays, const float* rays, float inv_scale,
                 pair_fn_t pair_fn, PencilCheckResult* out) {
    PencilCheckResult R;
    memset(&R, 0, sizeof(R));
    R.first_bad_ray = R.first_bad_tri = -1;
    PencilSetup S;
    memset(&S, 0, sizeof(S));
    bool ok;
    MirrorCheck MC;
    memset(&MC, 0, sizeof(MC));
    const bool mirror_mode = (mode == 3);
    if (mode == 0) ok = pencil_camera_setup(setup24, M_scene, nullptr, nullptr, S);
    else if (mode == 3) {
        // mode 3: the SHIPPED reflection pencil -- camera `setup24` mirrored about the plane of pencil_check_set_plane(); only rays
        // pencil_mirror_accepts() lets through are filtered (k_shade routes the others to the generic scan: counted as unsafe)
        PencilSetup cam;
        memset(&cam, 0, sizeof(cam));
        ok = pencil_camera_setup(setup24, M_scene, nullptr, nullptr, cam) && pencil_mirror_setup(cam, g_plane, g_plane[3], M_scene, nullptr, nullptr, S, MC);
        mode = 0;
    }
    else if (mode == 1) ok = pencil_light_setup(setup24, setup24 + 3, setup24 + 6, M_scene, S);
    else {
        double f[3] = {setup24[3], setup24[4], setup24[5]};
        const double fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        for (int k = 0; k < 3; ++k) { S.E[k] = setup24[k]; f[k] /= fl; }
        S.delta = setup24[6];
        S.w_max = setup24[7];
        const double lam_min = setup24[8];
        pencil_frame(S, f);
        pencil_finish_setup(S, M_scene);
        S.cos_g = std::fmax(kPencilCosMin, std::fmax(2.5 * S.delta / lam_min, 5.0 * S.theta));
        ok = fl > 0.0 && lam_min >= 2e-3 * S.M && S.w_max < 8.0;
        mode = 0;   // same ray semantics as primary rays from here on
    }
    R.setup_ok = ok ? 1 : 0;
    if (!ok) { *out = R; return 0; }
    R.delta = S.delta; R.M = S.M; R.cos_g = S.cos_g;

    std::vector<float> rec((size_t)ntri * 16);
    std::vector<uint8_t> state(ntri);   // 0 record, 1 always (not in the records), 2 never
    long long near_plane = 0;
    for (int i = 0; i < ntri; ++i) {
        const float *A = tri + 9 * i, *B = A + 3, *C = A + 6;
        float* q = &rec[(size_t)16 * i];
        // degenerate / NaN-barycentric triangles never reach the records (k_build_records)
        const float u[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, v[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
        const float n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        const float uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], uv = u[0] * v[0] + u[1] * v[1] + u[2] * v[2], vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        const float Df = uv * uv - uu * vv;
        if (n[0] == 0.f && n[1] == 0.f && n[2] == 0.f) { state[i] = 2; pencil_never(q); ++R.never_recs; continue; }
        const FilterTol t = filter_tolerances(A, B, C, dominant_axis(A, B, C), M_scene);
        if (t.always || !(std::fabs(Df) > 0.f) || !std::isfinite(Df)) { state[i] = 1; pencil_never(q); ++R.always_tris; continue; }
        if (!g_premise && pencil_plane_near(A, B, C, S)) { pencil_always(q); ++near_plane; continue; }   // as k_build_pencil does
        if (!pencil_record(A, B, C, t.E0, t.E1, S, q)) { state[i] = 2; ++R.never_recs; }
    }
    g_near_plane = near_plane;

    int64_t pairs = 0, ref_hits = 0, cands = 0, viol = 0, graz = 0, unsafe = 0;
    int bad_ray = -1, bad_tri = -1;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : pairs, ref_hits, cands, viol, graz, unsafe)
    for (int r = 0; r < nrays; ++r) {
        const float *O = rays + 6 * r, *D = O + 3;
        if (mirror_mode && !pencil_mirror_accepts(MC, O, D)) { ++unsafe; continue; }
        // pencil_set_slot (rt_kernels.cuh)
        const float sgn = mode == 1 ? -1.0f : 1.0f;
        const float dx = sgn * (D[0] - O[0]), dy = sgn * (D[1] - O[1]), dz = sgn * (D[2] - O[2]);
        const float den = fmaf(dx, S.F[6], fmaf(dy, S.F[7], dz * S.F[8]));
        const float inv = (1.0f / den) * inv_scale;
        float x = fmaf(dx, S.F[0], fmaf(dy, S.F[1], dz * S.F[2])) * inv;
        float y = fmaf(dx, S.F[3], fmaf(dy, S.F[4], dz * S.F[5])) * inv;
        const bool in_chart = (den > 0.0f) && (fmaf(x, x, fmaf(y, y, 1.0f)) <= S.w_max2);
        if (!in_chart) { ++unsafe; continue; }   // the kernels test these rays exactly against every triangle
        // pencil_zhi: depth of the origin + (nearest + s_lam) / |(x, y, 1)|, rounded up
        const float c = (1.0f / sqrtf(fmaf(x, x, fmaf(y, y, 1.0f)))) * 1.000002f * 1.0000005f;
        const float zO = fmaf(O[0] - S.Ef[0], S.F[6], fmaf(O[1] - S.Ef[1], S.F[7], (O[2] - S.Ef[2]) * S.F[8]));
        // true direction (double) for the grazing premise
        const double ddx = (double)D[0] - O[0], ddy = (double)D[1] - O[1], ddz = (double)D[2] - O[2];
        const double dl = std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        for (int i = 0; i < ntri; ++i) {
            ++pairs;
            const float* T = tri + 9 * i;
            float dist = 0.f;
            const int hit = pair_fn(O, D, T, T + 3, T + 6, &dist);
            const float* q = &rec[(size_t)16 * i];
            const float nearest = (mode == 0) ? (hit ? nextafterf(dist, INFINITY) : FLT_MAX) : 0.0f;
            const double zd = (double)zO + (double)round_up_sum(nearest, S.lam_slack) * (double)c;
            float zhi = (float)zd;
            if ((double)zhi < zd) zhi = nextafterf(zhi, INFINITY);
            if (!(zhi < FLT_MAX)) zhi = FLT_MAX;
            const float a = fmaf(q[0], x, fmaf(q[1], y, q[2]));
            const float b = fmaf(q[4], x, fmaf(q[5], y, q[6]));
            const float c2 = fmaf(q[8], x, fmaf(q[9], y, q[10]));
            const float sg = fmaf(q[3], x, fmaf(q[7], y, q[11]));
            const float e = fmaf(sg, zhi, q[12]);
            const float cc = c2;
            uint32_t ua, ub, uc, ue;
            memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); memcpy(&uc, &cc, 4); memcpy(&ue, &e, 4);
            const bool cand = !((ua | ub | uc | ue) >> 31);
            if (cand) ++cands;
            if (hit && dist < FLT_MAX) {
                ++ref_hits;
                if (state[i] == 1) continue;   // always-exact triangle: evaluated outside the filter
                if (!cand) {
                    // premise of the pencil kernels: pairs with |cos| < cos_g are certain misses in the reference
                    const float* A = T; const float* B = T + 3; const float* C = T + 6;
                    const double u[3] = {(double)B[0] - A[0], (double)B[1] - A[1], (double)B[2] - A[2]};
                    const double v[3] = {(double)C[0] - A[0], (double)C[1] - A[1], (double)C[2] - A[2]};
                    const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
                    const double cs = std::fabs(n[0] * ddx + n[1] * ddy + n[2] * ddz) / (std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) * dl);
                    if (g_premise && cs < S.cos_g) { ++graz; continue; }
                    ++viol;
#pragma omp critical
                    if (bad_ray < 0) { bad_ray = r; bad_tri = i; }
                }
            }
        }
    }
    R.pairs = pairs; R.ref_hits = ref_hits; R.candidates = cands; R.violations = viol; R.grazing_skipped = graz; R.unsafe_rays = unsafe;
    R.first_bad_ray = bad_ray; R.first_bad_tri = bad_tri;
    *out = R;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// The GENERIC filter (rt_kernels.cuh: k_build_records / fast_set / filter_pair), restated for the CPU.  Its only
// non-IEEE operations are MUFU.RCP and rsqrtf; `rc_scale` / `inv_scale` perturb them by a few ulp.  Same question as
// above: every pair the reference accepts must be a candidate -- with the tightest admissible distance bound in
// nearest-hit mode (mode 0: rhi from the next float above the pair's own distance), and with no bound (mode 1: any-hit).
// bmin: the grazing clause's threshold (1e-5), or < 0 for the clause-free records.
// ------------------------------------------------------------------------------------------------
int generic_check(int mode, double M, float bmin, int ntri, const float* tri, int nrays, const float* rays, float rc_scale, float inv_scale,
                  pair_fn_t pair_fn, PencilCheckResult* out) {
    PencilCheckResult R;
    memset(&R, 0, sizeof(R));
    R.first_bad_ray = R.first_bad_tri = -1;
    R.setup_ok = 1;
    R.M = M;
    const float cos_min = 1.0e-5f, u32 = 5.9604645e-8f;
    const float eps_r = (float)M * (48.0f * u32 / cos_min + 128.0f * u32);   // rt_b200.cu: eps_r_for
    std::vector<float> rec((size_t)ntri * 16);
    std::vector<uint8_t> state(ntri), cls(ntri);
    for (int i = 0; i < ntri; ++i) {
        const float *A = tri + 9 * i, *B = A + 3, *C = A + 6;
        float* q = &rec[(size_t)16 * i];
        for (int k = 0; k < 16; ++k) q[k] = 0.f;
        q[12] = -1.0f;   // "never"
        const float u[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, v[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
        const float n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        const float uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], uv = u[0] * v[0] + u[1] * v[1] + u[2] * v[2], vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        const float Df = uv * uv - uu * vv;
        if (n[0] == 0.f && n[1] == 0.f && n[2] == 0.f) { state[i] = 2; ++R.never_recs; continue; }
        const int W = dominant_axis(A, B, C);
        cls[i] = (uint8_t)W;
        const FilterTol t = filter_tolerances(A, B, C, W, M);
        if (t.always || !(std::fabs(Df) > 0.f) || !std::isfinite(Df)) { state[i] = 1; ++R.always_tris; continue; }
        const double a3[3] = {A[0], A[1], A[2]};
        const double inv = 1.0 / t.nn;
        const double n3[3] = {t.n3[0] * inv, t.n3[1] * inv, t.n3[2] * inv};
        q[0] = (float)n3[0]; q[1] = (float)n3[1]; q[2] = (float)n3[2]; q[3] = (float)(-(n3[0] * a3[0] + n3[1] * a3[1] + n3[2] * a3[2]));
        q[4] = (float)t.su; q[5] = (float)t.sv; q[6] = (float)(-(t.su * a3[t.U] + t.sv * a3[t.V]) + t.E0); q[7] = (float)(1.0 + 3.0 * t.E0);
        q[8] = (float)t.tu; q[9] = (float)t.tv; q[10] = (float)(-(t.tu * a3[t.U] + t.tv * a3[t.V]) + t.E0); q[11] = (float)(-t.E1);
        q[12] = bmin;
    }
    int64_t pairs = 0, ref_hits = 0, cands = 0, viol = 0;
    int bad_ray = -1, bad_tri = -1;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : pairs, ref_hits, cands, viol)
    for (int r = 0; r < nrays; ++r) {
        const float *O = rays + 6 * r, *D = O + 3;
        // fast_set
        float d[3] = {D[0] - O[0], D[1] - O[1], D[2] - O[2]};
        const float len2 = fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0]));
        float inv = (1.0f / sqrtf(len2)) * inv_scale;
        if (!(len2 > 1e-30f) || !(len2 < 1e30f)) inv = 0.0f;
        for (int k = 0; k < 3; ++k) d[k] *= inv;
        const float o[3] = {fmaf(-eps_r, d[0], O[0]), fmaf(-eps_r, d[1], O[1]), fmaf(-eps_r, d[2], O[2])};
        for (int i = 0; i < ntri; ++i) {
            ++pairs;
            const float* T = tri + 9 * i;
            float dist = 0.f;
            const int hit = pair_fn(O, D, T, T + 3, T + 6, &dist);
            const float* q = &rec[(size_t)16 * i];
            const int W = cls[i], U = (W + 1) % 3, V = (W + 2) % 3;
            // filter_pair
            float b = q[0] * d[0];
            b = fmaf(q[1], d[1], b);
            b = fmaf(q[2], d[2], b);
            float a = fmaf(q[0], o[0], q[3]);
            a = fmaf(q[1], o[1], a);
            a = fmaf(q[2], o[2], a);
            const float rc = (1.0f / (-b)) * rc_scale;
            const float rr = a * rc;
            const float iu = fmaf(rr, d[U], o[U]), iv = fmaf(rr, d[V], o[V]);
            float s = fmaf(q[4], iu, q[6]);
            s = fmaf(q[5], iv, s);
            float t = fmaf(q[8], iu, q[10]);
            t = fmaf(q[9], iv, t);
            float qq = q[7] + (-s);
            qq = qq + (-t);
            const float m = fminf(fminf(s, t), qq);
            const float e = q[11] * rc;
            uint32_t rbits, rhi = 0x7f7fffffu;
            memcpy(&rbits, &rr, 4);
            if (mode == 0 && hit) {
                const float best = nextafterf(dist, INFINITY);
                const float hi = round_up_sum(best, 2.0f * eps_r);
                memcpy(&rhi, &hi, 4);
            }
            bool cand = !(m < -std::fabs(e)) && (rbits < rhi);
            if (bmin > 0.0f) cand = cand || (std::fabs(b) < bmin);
            if (cand) ++cands;
            if (hit && dist < FLT_MAX) {
                ++ref_hits;
                if (state[i] == 1) continue;
                if (!cand) {
                    ++viol;
#pragma omp critical
                    if (bad_ray < 0) { bad_ray = r; bad_tri = i; }
                }
            }
        }
    }
    R.pairs = pairs; R.ref_hits = ref_hits; R.candidates = cands; R.violations = viol;
    R.first_bad_ray = bad_ray; R.first_bad_tri = bad_tri;
    *out = R;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Thread pencils (rt_tpencil.h): continuation rays grouped by the triangle their primary ray hit; the rays of a group share
// the mirror image of `eye` about that triangle's plane.  Replays what k_tp_route / k_trace_tp do: tp_mirror_point,
// tp_accepts (a refused ray goes to the generic scan: counted as unsafe), groups of R accepted rays of one reflector in
// input order, tp_thread_consts from the group's smallest lam_o, and for EVERY (ray, triangle) pair tp_orient + the hot
// test (tp_weights / near) + the cold path's full test (tp_candidate) with the tightest admissible "nearest so far".
// ray_tri[r]: the reflector of ray r.  out->candidates counts the HOT candidates (what enters the cold path),
// out->grazing_skipped the pairs that pass the full test with no distance bound (what would be evaluated exactly at most).
// ------------------------------------------------------------------------------------------------
int tpencil_check(const double* eye3, double delta_cam, double M_scene, const float* box_lo, const float* box_hi, int ntri, const float* tri, int nrays,
                  const float* rays, const int32_t* ray_tri, int R, pair_fn_t pair_fn, PencilCheckResult* out) {
    PencilCheckResult res;
    memset(&res, 0, sizeof(res));
    res.first_bad_ray = res.first_bad_tri = -1;
    TpSetup S;
    memset(&S, 0, sizeof(S));
    const bool ok = tp_setup(eye3, delta_cam, M_scene, box_lo, box_hi, S);
    res.setup_ok = ok ? 1 : 0;
    if (!ok) { *out = res; return 0; }
    res.delta = S.delta; res.M = S.M; res.cos_g = S.cg_floor;
    std::vector<float> rec((size_t)ntri * 24);
    std::vector<uint8_t> state(ntri);   // 0 record, 1 always-exact (outside the filter), 2 never
    for (int i = 0; i < ntri; ++i) {
        const float *A = tri + 9 * i, *B = A + 3, *C = A + 6;
        float* q = &rec[(size_t)24 * i];
        const float u[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, v[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
        const float n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        const float uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], uv = u[0] * v[0] + u[1] * v[1] + u[2] * v[2], vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        const float Df = uv * uv - uu * vv;
        if (n[0] == 0.f && n[1] == 0.f && n[2] == 0.f) { state[i] = 2; tp_never(q); ++res.never_recs; continue; }
        const FilterTol t = filter_tolerances(A, B, C, dominant_axis(A, B, C), M_scene);
        if (t.always || !(std::fabs(Df) > 0.f) || !std::isfinite(Df)) { state[i] = 1; tp_never(q); ++res.always_tris; continue; }
        if (!tp_record(A, B, C, t.E0, t.E1, S, q)) { state[i] = 2; ++res.never_recs; }
    }
    // accepted rays per reflector, in input order
    std::vector<std::vector<int>> by_tri(ntri);
    std::vector<TpRay> tr(nrays);
    std::vector<float> Es((size_t)ntri * 3, 0.f);
    std::vector<uint8_t> has_E(ntri, 0);
    int64_t unsafe = 0;
    for (int r = 0; r < nrays; ++r) {
        const int t = ray_tri[r];
        if (t < 0 || t >= ntri) { ++unsafe; continue; }
        const float* T = tri + 9 * t;
        if (!has_E[t]) has_E[t] = tp_mirror_point(S.eye, S.centerf, T, T + 3, T + 6, &Es[(size_t)3 * t]) ? 1 : 2;
        if (has_E[t] != 1 || !tp_accepts(S, &Es[(size_t)3 * t], rays + 6 * r, rays + 6 * r + 3, tr[r])) { ++unsafe; continue; }
        by_tri[t].push_back(r);
    }
    std::vector<std::pair<int, int>> groups;   // (reflector, first index into by_tri[reflector])
    for (int t = 0; t < ntri; ++t)
        for (size_t g0 = 0; g0 < by_tri[t].size(); g0 += (size_t)R) groups.emplace_back(t, (int)g0);
    int64_t pairs = 0, ref_hits = 0, cands = 0, viol = 0, full = 0;
    int bad_ray = -1, bad_tri = -1;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : pairs, ref_hits, cands, viol, full)
    for (size_t g = 0; g < groups.size(); ++g) {
        const int t = groups[g].first;
        const std::vector<int>& list = by_tri[t];
        const size_t g0 = (size_t)groups[g].second, g1 = std::min(list.size(), g0 + (size_t)R);
        const float* E = &Es[(size_t)3 * t];
        float lam_min_thread = FLT_MAX;
        for (size_t k = g0; k < g1; ++k) lam_min_thread = std::fmin(lam_min_thread, tr[list[k]].lam_o);
        float cg, near_thr;
        tp_thread_consts(S, lam_min_thread, cg, near_thr);
        for (int i = 0; i < ntri; ++i) {
            TpTri T;
            tp_orient(&rec[(size_t)24 * i], E, near_thr, T);
            const float* V = tri + 9 * i;
            for (size_t k = g0; k < g1; ++k) {
                const int r = list[k];
                const float *O = rays + 6 * r, *D = O + 3;
                ++pairs;
                float dist = 0.f;
                const int hit = pair_fn(O, D, V, V + 3, V + 6, &dist);
                const bool hot = !(tp_weights(T, tr[r]) >> 31) || T.near_;
                if (hot) ++cands;
                const float lam_lo = nextafterf(tr[r].lam_o - S.lam_slack, -INFINITY);
                if (tp_candidate(T, tr[r], FLT_MAX, lam_lo, cg)) ++full;
                if (hit && dist < FLT_MAX) {
                    ++ref_hits;
                    if (state[i] == 1) continue;   // always-exact triangle: evaluated outside the filter
                    const float nearest = nextafterf(dist, INFINITY);
                    float lam_hi = round_up_sum(tr[r].lam_o, round_up_sum(nearest, S.lam_slack));
                    if (!(lam_hi < FLT_MAX)) lam_hi = FLT_MAX;
                    if (!hot || !tp_candidate(T, tr[r], lam_hi, lam_lo, cg)) {
                        ++viol;
#pragma omp critical
                        if (bad_ray < 0) { bad_ray = r; bad_tri = i; }
                    }
                }
            }
        }
    }
    res.pairs = pairs; res.ref_hits = ref_hits; res.candidates = cands; res.violations = viol; res.grazing_skipped = full; res.unsafe_rays = unsafe;
    res.first_bad_ray = bad_ray; res.first_bad_tri = bad_tri;
    *out = res;
    return 0;
}

}  // extern "C"
