// pencil_check.cpp -- TEST INFRASTRUCTURE: CPU replay of the pencil filter (raytracert_b200/csrc/rt_pencil.h).
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

extern "C" {

// mode 0: primary rays of the camera `setup24` (24 corner floats); mode 1: shadow rays ending at the light `setup24[0..2]`,
// box = setup24[3..5] (lo) / [6..8] (hi).  tri: ntri x 9 floats.  rays: n x 6 floats (origin, dest).
// inv_scale perturbs the normalisation factor (emulates the 2-ulp error of rsqrtf): w = dir * (1/sqrt(len2)) * inv_scale.
int pencil_check(int mode, const float* setup24, double M_scene, int ntri, const float* tri, int nrays, const float* rays, float inv_scale,
                 pair_fn_t pair_fn, PencilCheckResult* out) {
    PencilCheckResult R;
    memset(&R, 0, sizeof(R));
    R.first_bad_ray = R.first_bad_tri = -1;
    PencilSetup S;
    int axis = 0;
    float sign = 0.f;
    bool ok;
    if (mode == 0) ok = pencil_camera_setup(setup24, M_scene, S);
    else ok = pencil_light_setup(setup24, setup24 + 3, setup24 + 6, M_scene, S, axis, sign);
    R.setup_ok = ok ? 1 : 0;
    if (!ok) { *out = R; return 0; }
    R.delta = S.delta; R.M = S.M; R.cos_g = S.cos_g;

    std::vector<float> rec((size_t)ntri * 16);
    std::vector<uint8_t> state(ntri);   // 0 record, 1 always (not in the records), 2 never
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
        if (!pencil_record(A, B, C, t.E0, t.E1, S, q)) { state[i] = 2; ++R.never_recs; }
    }

    int64_t pairs = 0, ref_hits = 0, cands = 0, viol = 0, graz = 0, unsafe = 0;
    int bad_ray = -1, bad_tri = -1;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : pairs, ref_hits, cands, viol, graz, unsafe)
    for (int r = 0; r < nrays; ++r) {
        const float *O = rays + 6 * r, *D = O + 3;
        if (mode == 1 && !(sign * (D[axis] - O[axis]) >= 0.0f)) { ++unsafe; continue; }   // the kernel tests these exactly
        // pencil_set_slot
        float dx = D[0] - O[0], dy = D[1] - O[1], dz = D[2] - O[2];
        const float len2 = dx * dx + dy * dy + dz * dz;
        float inv = (1.0f / sqrtf(len2)) * inv_scale;
        const bool degenerate = !(len2 > 1e-30f) || !(len2 < 1e30f);
        if (mode == 1) inv = -inv;
        dx *= inv; dy *= inv; dz *= inv;
        float lam = fmaf(dx, O[0] - S.Ef[0], fmaf(dy, O[1] - S.Ef[1], dz * (O[2] - S.Ef[2])));
        if (degenerate) { dx = dy = dz = 0.f; }
        // true direction (double) for the grazing premise
        const double ddx = (double)D[0] - O[0], ddy = (double)D[1] - O[1], ddz = (double)D[2] - O[2];
        const double dl = std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        for (int i = 0; i < ntri; ++i) {
            ++pairs;
            const float* T = tri + 9 * i;
            float dist = 0.f;
            const int hit = pair_fn(O, D, T, T + 3, T + 6, &dist);
            const float* q = &rec[(size_t)16 * i];
            float lhi;
            if (degenerate) lhi = INFINITY;
            else if (mode == 0) {
                const float best = hit ? nextafterf(dist, INFINITY) : FLT_MAX;
                lhi = round_up_sum(round_up_sum(lam, best), S.lam_slack);
                if (!(lhi < FLT_MAX)) lhi = FLT_MAX;
            } else {
                lhi = round_up_sum(round_up_sum(lam, 0.0f), S.lam_slack);
                if (!(lhi < FLT_MAX)) lhi = FLT_MAX;
            }
            const float a = fmaf(q[0], dx, fmaf(q[1], dy, fmaf(q[2], dz, q[3])));
            const float b = fmaf(q[4], dx, fmaf(q[5], dy, fmaf(q[6], dz, q[7])));
            const float c = fmaf(q[8], dx, fmaf(q[9], dy, fmaf(q[10], dz, q[11])));
            const float sg = (a + b) + c;
            const float e = fmaf(sg, lhi, q[12]);
            uint32_t ua, ub, uc, ue;
            memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); memcpy(&uc, &c, 4); memcpy(&ue, &e, 4);
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
                    if (cs < S.cos_g) { ++graz; continue; }
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

}  // extern "C"
