// rt_kernels.cuh -- the sm_100a kernels of the render hot path.
//
// Reference functions replaced (paths under /root/reference/CG_Project):
//   k_build_records / k_build_tile_boxes -- (new) per-triangle filter records, built from u,v,n,uu,uv,vv,D of
//                        raytracing.cpp:106-140, and per-tile bounds for the opt-in tile culling
//   k_build_pencil    -- (new) per-frame records of the common-point ("pencil") filter around the eye / a light (rt_pencil.h)
//   k_trace           -- main.cpp:377-388 ray generation (PRIMARY) + intersectMesh raytracing.cpp:161-192
//                        (PENCIL: primary rays through the common-point filter)
//   k_finish          -- tail of intersectMesh / head of trace (raytracing.cpp:183-191, 387-396): winner's hit point,
//                        analytic spheres, hit record
//   k_shadow          -- isShadow raytracing.cpp:241-261 (PENCIL: any-hit rays of one light through the common-point filter)
//   k_shade           -- shade/diffuseOnly/blinnPhongSpecularOnly/reflection/refraction/addOffset/trace
//                        raytracing.cpp:197-232, 266-330, 335-406
//   k_trace<..,!PRIMARY,PENCIL> -- the same scan for the level-1 continuation rays of a plane group around the mirrored eye
//                        (reflection(), raytracing.cpp:277-285, mirrors a pencil; rt_pencil.h: pencil_mirror_setup)
//   k_build_trec / k_tp_offsets / k_tp_scatter / k_trace_tp -- thread pencils (rt_tpencil.h): level-1 continuation rays grouped by the
//                        triangle their primary ray hit, 8 per thread, weights built on the fly around that triangle's mirror
//                        image of the eye
//   k_trace_small     -- performRayTracing raytracing.cpp:410-416 for a handful of rays: the whole recursion in one launch
//   k_init_trace      -- rt_trace: unpacks the uploaded (origin, dest) pairs
//   k_resolve         -- main.cpp:391-393 + RGBValue clamp main.cpp:24-42
//   k_deinterleave / k_place_rows / k_quantise -- row gather after the all-gather; Image::writeImage's quantiser main.cpp:117
//
// How parity and speed coexist (DESIGN.md "filter + exact"):
//   every (ray, triangle) pair first goes through a CONSERVATIVE FILTER evaluated with packed FP32 FMAs
//   (FFMA2, two rays per instruction) on a precomputed 64-byte record; the filter may only say
//   "certainly not a hit / certainly not nearer than the current best".  Pairs it cannot rule out are
//   re-evaluated by exact_ray_triangle(), the reference's expression order with non-contracted IEEE
//   operations, and only that exact result ever updates the nearest hit -- so primitive ids and hit
//   points are the reference's, bit for bit, while ~all the work runs at FMA-pipe speed.
//
// Work decomposition: a scan launch is cut into items = (chunk of kThreads*R rays) x (range of triangle tiles);
// persistent CTAs stride over the items, tiles stream through a TMA-filled shared-memory ring, per-ray results of
// the parts are merged with atomics (make_split / Pipe / scan_pass below).
#pragma once
#include "rt_common.cuh"
#include "rt_pencil.h"
#include "rt_tpencil.h"
#include "../../include/rt_b200.h"

namespace rt {

constexpr int kTile = 128;         // triangles per shared-memory tile
constexpr int kRecVec = 4;         // float4 per filter record (64 B)
constexpr int kStages = 4;         // TMA ring depth
constexpr int kWarps = 8;          // all warps are consumers; the last warp to finish a stage refills it
constexpr int kThreads = kWarps * 32;
constexpr int kMaxParts = 64;      // a launch with few rays splits the triangle range into <= kMaxParts parts per ray chunk
constexpr uint32_t kItemsPerCta = 24;   // load-balance target of make_split
constexpr uint32_t kMinPartTiles = 8;   // >= 1024 triangles per part unless the launch is tiny
constexpr int kPadTiles = kMaxParts;  // "never" tiles appended to the record array so that every part has the same length
constexpr int kSuper = 32;              // tiles per super-tile (hierarchical tile culling)
constexpr int kCullMaxTiles = 16384;    // the CTA-level "needed tiles" bitmap covers this many tiles (2.1 M triangles)
constexpr int kCullList = 4096;         // needed tiles streamed per batch
constexpr unsigned long long kKeyEmpty = ~0ull;  // (distance bits << 32 | primitive id); all ones = no hit yet

constexpr float kU32 = 5.9604645e-8f;  // 2^-24
constexpr float kCosMinDefault = 1.0e-5f;  // |cos(ray, plane normal)| below this -> always exact ("grazing")

// counters[] layout (uint32 unless noted)
constexpr int kMaxLevels = 40;
constexpr int kCntHit = 0;                 // [level] hits recorded by k_finish at that level
constexpr int kCntRay = kMaxLevels;        // [level] rays queued for k_trace at that level
constexpr int kCntExact = 2 * kMaxLevels;  // 64-bit: exact re-evaluations (2 words)
constexpr int kMaxMirrors = 4;             // reflection pencils: plane groups served per frame (rt_pencil.h: pencil_mirror_setup)
constexpr int kCntMirror = 2 * kMaxLevels + 4;   // [group] level-1 continuation rays routed to the mirror pencil of that plane group
constexpr int kCntTpPool = 2 * kMaxLevels + 4 + kMaxMirrors;       // thread pencils: level-1 continuation rays accepted into the pool
constexpr int kCntTpGroups = kCntTpPool + 1;                      // ... and the groups of kTpR rays formed from them
constexpr int kCntWords = 2 * kMaxLevels + 4 + kMaxMirrors + 2;
constexpr int kTpR = 8;                    // rays per thread of the thread-pencil scan
constexpr int kQueueTp = 100;              // FrameParams::mirror_sel value that selects the thread-pencil queue
constexpr uint8_t kNoGroup = 0xff;

struct FrameParams {
    // scene
    const float4* rec;          // filter records, ntiles*kTile*4 float4
    const float4* triv;         // exact corners: 3 float4 per triangle
    const float4* normal_mat;   // xyz unit face normal, w = material index (bits)
    const float4* materials;    // 4 float4 per material: Kd|Ns, Ka|Ni, Ks|Tr, flags
    const float4* spheres;      // 2 float4 per sphere: center|radius, material bits
    int ntri, ntiles, nspheres;
    int cls1, cls2;             // first tile of dominant-axis class 1 / class 2 (records are grouped by class)
    // per-sample state (chunk local)
    float4* ray_o;              // xyz origin
    float4* ray_d;              // xyz dest, w = lvl (bits)
    float4* thr;                // rgb throughput (product of K along the chain)
    float4* acc;                // rgb accumulated colour
    float4* hit;                // xyz intersection, w = primitive id (bits), -1 = miss
    uint32_t* lit;              // bit l: light l reaches the hit point
    uint32_t* q_ray;
    uint32_t* q_hit;
    uint32_t* counters;
    unsigned long long* key;    // per sample: nearest (distance, triangle) found by the scan parts, merged with atomicMin
    const uint32_t* always_list;  // triangles the filter cannot bound ("always exact"): evaluated per ray outside the scan
    int n_always;
    const float4* tile_box;     // 2 float4 per tile: conservative AABB (lo, hi) of the tile's candidate region
    int cull;                   // RT_OPT_TILE_CULLING in effect for this launch (the CULL = true kernel instantiations)
    const float4* super_box;    // 2 float4 per super-tile (kSuper tiles)
    int32_t* prim_out;          // optional: primary primitive id per local sample
    // frame
    float corners[24];
    float divX, divY;
    uint32_t W, H, pfx, pfy;
    uint32_t row0, nrows;       // local rows of this chunk
    uint32_t G, rank;           // row interleave: global y = local_row * G + rank
    uint32_t nsamples;          // samples in this chunk
    uint32_t nslots;            // primary work slots: samples in tiled order, padded to whole 8x8-pixel blocks (== nsamples for rt_trace)
    uint32_t tiles_x;           // 8x8-pixel blocks per row of blocks
    unsigned long long sample_base;  // local sample index of the chunk's first sample (prim_out addressing)
    float eps_r;                // distance guard band of the filter
    float camera[3];
    int nlights;
    float lights[RT_MAX_LIGHTS][3];
    uint32_t features;
    int max_lvl;
    int trace_api;              // 1: rays come from ray_o/ray_d (rt_trace), not from the camera
    // pencil launches (rt_pencil.h): every ray's line passes through the common point pE
    const float4* prec;         // pencil records of this launch (same positions / tiles as rec); nullptr: generic filter
    float pE[3];                // the common point, float
    float p_lam_slack;          // s_lam: ray-side guard band of the distance clause
    int light_sel;              // k_shadow: >= 0 -> this launch handles only that light; -1 -> all lights (ray = hit * nlights + light)
    float pF[9];                // chart frame u, v, f: a ray is (x, y, 1) = dir / (dir.f) in it
    float p_wmax2;              // chart bound: 1 + x^2 + y^2 <= p_wmax2, else the ray is not filtered (exact path for everything)
    // reflection pencils (level-1 continuation rays of primary hits on a plane group; rt_pencil.h)
    int n_mirrors;              // plane groups with a mirror pencil this frame (0: none)
    const uint8_t* tri_group;   // plane group of every triangle (kNoGroup: none); only read when n_mirrors > 0
    uint32_t* q_mirror;         // ray queues of the groups, q_mirror_stride entries each
    uint32_t q_mirror_stride;
    int mirror_sel;             // >= 0: this k_trace / k_finish launch serves the ray queue of that group; -1: the ordinary queue
    MirrorCheck mirror[kMaxMirrors];
    // thread pencils (rt_tpencil.h): level-1 continuation rays grouped by the triangle their primary ray hit
    int tp_on;                  // k_shade (level 0) offers eligible continuation rays to the pool
    const float4* trec;         // thread-pencil records: kTpVec float4 per record position (same positions / tiles as rec)
    uint32_t* tp_pool;          // accepted rays (sample ids), unsorted
    uint32_t* tp_hist;          // [ntri] accepted rays per reflector
    uint32_t* tp_off;           // [ntri] first group of the reflector
    uint32_t* tp_cursor;        // [ntri]
    uint32_t* q_tp;             // grouped queue: kTpR consecutive entries share a reflector
    uint32_t* tp_group_tri;     // reflector of every group
    TpSetup tp;
};
__device__ __forceinline__ const uint32_t* ray_queue(const FrameParams& P) {
    if (P.mirror_sel == kQueueTp) return P.q_tp;
    return P.mirror_sel >= 0 ? P.q_mirror + (size_t)P.mirror_sel * P.q_mirror_stride : P.q_ray;
}
__device__ __forceinline__ uint32_t ray_queue_count(const FrameParams& P, int level) {
    if (P.mirror_sel == kQueueTp) return P.counters[kCntTpGroups] * (uint32_t)kTpR;
    return P.mirror_sel >= 0 ? P.counters[kCntMirror + P.mirror_sel] : P.counters[kCntRay + level];
}

// ------------------------------------------------------------------------------------------------
// Filter records
// ------------------------------------------------------------------------------------------------
// Triangles are grouped by the dominant axis W of their plane normal (class 0: W = x, 1: W = y, 2: W = z; `perm`
// maps record position -> triangle id, classes are padded to whole tiles with "never" records).  On a plane that is not
// parallel to W the barycentrics are affine functions of the two other coordinates (U, V) alone, so the filter needs
// only two components of the plane hit point and two-term functionals: 16 packed FP32 instructions per (ray pair,
// triangle) instead of 19.  (U, V) = (y, z), (z, x), (x, y) for W = x, y, z.
//
// Record at position p (4 float4):
//   q0 = ( nx, ny, nz, dn )     unit plane normal, dn = -n.T0        -> h = n.O' + dn, cos = n.d
//   q1 = ( su, sv, cs', c1 )    s(P) = su*Pu + sv*Pv + cs, cs' = cs + E0   (first barycentric of raytracing.cpp:144), c1 = 1 + 3*E0
//   q2 = ( tu, tv, ct', -E1 )   t(P) = tu*Pu + tv*Pv + ct, ct' = ct + E0   (second barycentric, :148)
//   q3 = ( bmin, id, nv, 0 )    id = triangle index (bits); nv (first record of a tile only) = records in use in the tile
// A pair is a CANDIDATE (goes to the exact path) iff
//   ( min(s', t', c1 - s' - t') >= -E1*|1/cos|  and  0 <= r' < rhi' )  or  |cos| < bmin
// with r' = h/(-cos) (distance along the unit direction from the shifted origin O' = O - eps_r*d).
//   bmin = cos_min : normal triangle;   bmin = -1 : never a candidate (degenerate n == 0, padding);
//   (triangles the filter cannot bound are not in the records at all: see always_list below)
// E0/E1 bound the difference between this evaluation and the reference's own rounding.
constexpr uint32_t kNoTriangle = 0xffffffffu;

// bmin_regular: cos_min, or kBminNoGrazing (< 0, never true) when the grazing clause is provably unnecessary.
// Triangles the filter cannot bound (float D == 0 / NaN so that the reference's NaN barycentrics pass its tests,
// non-finite vertices, slivers with E0 >= 64) get a "never" record and are appended to always_list instead: k_finish /
// k_shadow evaluate those exactly for every ray.  Nearest hits merge by (distance, id) and shadow rays only ask
// "any hit", so where a triangle is evaluated does not matter.
constexpr float kBminNever = -1.0f, kBminNoGrazing = -0.5f;
__global__ void k_build_records(const float4* __restrict__ triv, const uint32_t* __restrict__ perm, int npos, int c1_end, int c2_end, float M,
                                float bmin_regular, float4* __restrict__ rec, unsigned int* __restrict__ n_always, uint32_t* __restrict__ always_list) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const uint32_t i = perm[pos];
    const int W = pos < c1_end ? 0 : (pos < c2_end ? 1 : 2);   // class of this position (tile granular)
    float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = make_float4(kBminNever, __uint_as_float(i), 0, 0);  // "never"
    if (i != kNoTriangle) {
        const float4 A = triv[3 * i], B = triv[3 * i + 1], C = triv[3 * i + 2];
        // the reference's own float quantities decide degeneracy (raytracing.cpp:106-109,134-140)
        v3 uf = e_sub(mk3(B), mk3(A)), vf = e_sub(mk3(C), mk3(A));
        v3 nf = e_cross(uf, vf);
        const bool null_n = (nf.x == 0.0f && nf.y == 0.0f && nf.z == 0.0f);
        float uuf = e_dot(uf, uf), uvf = e_dot(uf, vf), vvf = e_dot(vf, vf);
        float Df = __fsub_rn(__fmul_rn(uvf, uvf), __fmul_rn(uuf, vvf));
        if (!null_n) {
            bool always = !(fabsf(Df) > 0.0f) || !isfinite(Df);
            const float a3f[3] = {A.x, A.y, A.z}, b3f[3] = {B.x, B.y, B.z}, c3f[3] = {C.x, C.y, C.z};
            const FilterTol t = filter_tolerances(a3f, b3f, c3f, W, (double)M);   // rt_pencil.h: shared with the pencil records
            if (t.always) always = true;
            if (!always) {
                const double a3[3] = {A.x, A.y, A.z};
                const double inv = 1.0 / t.nn;
                const double n3[3] = {t.n3[0] * inv, t.n3[1] * inv, t.n3[2] * inv};
                q0 = make_float4((float)n3[0], (float)n3[1], (float)n3[2], (float)(-(n3[0] * a3[0] + n3[1] * a3[1] + n3[2] * a3[2])));
                q1 = make_float4((float)t.su, (float)t.sv, (float)(-(t.su * a3[t.U] + t.sv * a3[t.V]) + t.E0), (float)(1.0 + 3.0 * t.E0));
                q2 = make_float4((float)t.tu, (float)t.tv, (float)(-(t.tu * a3[t.U] + t.tv * a3[t.V]) + t.E0), (float)(-t.E1));
                q3.x = bmin_regular;
            }
            if (always) { q0 = q1 = q2 = make_float4(0, 0, 0, 0); q3.x = kBminNever; always_list[atomicAdd(n_always, 1u)] = i; }
        }
    }
    if ((pos % kTile) == 0) {
        // first record of a tile: q3.z = number of records up to the tile's last triangle (padding sits at the end of a
        // class, so a partially filled tile -- every tile of a tiny scene -- is scanned only that far)
        int last = -1;
        for (int j = 0; j < kTile; ++j) if (perm[pos + j] != kNoTriangle) last = j;
        q3.z = __int_as_float(last + 1);
    }
    rec[4 * pos] = q0; rec[4 * pos + 1] = q1; rec[4 * pos + 2] = q2; rec[4 * pos + 3] = q3;
}

// Conservative bounding box of everything a tile can make a hit of (RT_OPT_TILE_CULLING).
// A hit accepted by the exact path has its point I within 8uM of the ray, within 4uM of the triangle's plane and,
// because the reference's own barycentrics are off by at most E0 (DESIGN.md "filter soundness"), inside the triangle
// dilated by E0 in barycentric units, i.e. by <= 3*E0*diameter.  So a ray that misses the box of the dilated
// triangles (+ rounding slack) cannot hit any of them.  E0 is read back from the record (c1 = 1 + 3*E0).
// "never" records (degenerate, padding, and the triangles that went to always_list) contribute nothing; a record
// with a non-finite bmin (not produced any more) would make its tile unbounded, i.e. never skipped.
__global__ void k_build_tile_boxes(const float4* __restrict__ triv, const float4* __restrict__ rec, int ntiles_padded, float M,
                                   float4* __restrict__ tile_box) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntiles_padded) return;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    bool unbounded = false;
    for (int j = 0; j < kTile; ++j) {
        const int pos = tile * kTile + j;
        const float4 q3 = rec[4 * pos + 3];
        if (q3.x == kBminNever) continue;                   // never a candidate
        if (!(q3.x < inf)) { unbounded = true; break; }     // defensive: no bound
        const uint32_t i = __float_as_uint(q3.y);
        const float4 A = triv[3 * i], B = triv[3 * i + 1], C = triv[3 * i + 2];
        const float e0 = fmaxf(rec[4 * pos + 1].w - 1.0f, 0.0f) * (1.0f / 3.0f) + 1e-6f;
        const float ab = sqrtf((B.x - A.x) * (B.x - A.x) + (B.y - A.y) * (B.y - A.y) + (B.z - A.z) * (B.z - A.z));
        const float ac = sqrtf((C.x - A.x) * (C.x - A.x) + (C.y - A.y) * (C.y - A.y) + (C.z - A.z) * (C.z - A.z));
        const float bc = sqrtf((C.x - B.x) * (C.x - B.x) + (C.y - B.y) * (C.y - B.y) + (C.z - B.z) * (C.z - B.z));
        const float m = 4.0f * e0 * fmaxf(ab, fmaxf(ac, bc)) + M * 6.103515625e-5f;   // dilation (x4/3 slack) + 2^-14 M >> 12uM
        const float4 v[3] = {A, B, C};
        for (int k = 0; k < 3; ++k) {
            const float c[3] = {v[k].x, v[k].y, v[k].z};
            for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], c[a] - m); hi[a] = fmaxf(hi[a], c[a] + m); }
        }
    }
    if (unbounded) { for (int a = 0; a < 3; ++a) { lo[a] = -inf; hi[a] = inf; } }
    // an empty tile keeps lo = +inf, hi = -inf: no ray reaches it
    tile_box[2 * tile] = make_float4(lo[0], lo[1], lo[2], 0.f);
    tile_box[2 * tile + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
}

// Super-tile boxes: union of kSuper consecutive tile boxes (tiles are Morton-sorted when the option was set before the upload).
__global__ void k_build_super_boxes(const float4* __restrict__ tile_box, int ntiles_padded, int nsuper, float4* __restrict__ super_box) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nsuper) return;
    const float inf = __int_as_float(0x7f800000);
    float4 lo = make_float4(inf, inf, inf, 0.f), hi = make_float4(-inf, -inf, -inf, 0.f);
    for (int t = g * kSuper; t < min((g + 1) * kSuper, ntiles_padded); ++t) {
        const float4 a = tile_box[2 * t], b = tile_box[2 * t + 1];
        lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
        hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
    }
    super_box[2 * g] = lo; super_box[2 * g + 1] = hi;
}

// Union of the tile boxes = bounding box of everything a filter record can make a hit of (pencil_light_setup).
__global__ void k_scene_box(const float4* __restrict__ super_box, int nsuper, float4* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const float inf = __int_as_float(0x7f800000);
    float4 lo = make_float4(inf, inf, inf, 0.f), hi = make_float4(-inf, -inf, -inf, 0.f);
    for (int g = 0; g < nsuper; ++g) {
        const float4 a = super_box[2 * g], b = super_box[2 * g + 1];
        lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
        hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
    }
    out[0] = lo; out[1] = hi;
}

// Pencil records around the common point S.E (rt_pencil.h), position by position next to the generic records: a
// position that is "never" there (padding, degenerate, always-exact triangle) is "never" here; id and the tile's
// record count are copied.  M is the magnitude bound the generic records were built with (same E0 / E1).
// premise = 0 (RT_OPT_PENCIL_ANY, no clause-free proof): triangles whose plane passes too close to the common point get an
// "always candidate" record and are counted in *n_near (rt_pencil.h: pencil_plane_near).
__global__ void k_build_pencil(const float4* __restrict__ triv, const float4* __restrict__ rec, int npos, int c1_end, int c2_end, float M,
                               const PencilSetup S, float4* __restrict__ prec, int premise, unsigned int* __restrict__ n_near) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const float4 g3 = rec[4 * pos + 3];
    float q[16];
    pencil_never(q);
    if (g3.x != kBminNever) {
        const uint32_t i = __float_as_uint(g3.y);
        const int W = pos < c1_end ? 0 : (pos < c2_end ? 1 : 2);
        const float4 A = triv[3 * i], B = triv[3 * i + 1], C = triv[3 * i + 2];
        const float a3f[3] = {A.x, A.y, A.z}, b3f[3] = {B.x, B.y, B.z}, c3f[3] = {C.x, C.y, C.z};
        const FilterTol t = filter_tolerances(a3f, b3f, c3f, W, (double)M);
        if (!t.always) {
            if (!premise && pencil_plane_near(a3f, b3f, c3f, S)) { pencil_always(q); atomicAdd(n_near, 1u); }
            else pencil_record(a3f, b3f, c3f, t.E0, t.E1, S, q);
        }
    }
    prec[4 * pos] = make_float4(q[0], q[1], q[2], q[3]);
    prec[4 * pos + 1] = make_float4(q[4], q[5], q[6], q[7]);
    prec[4 * pos + 2] = make_float4(q[8], q[9], q[10], q[11]);
    prec[4 * pos + 3] = make_float4(q[12], g3.y, g3.z, 0.f);
}

// Slab test of the half-line O' + t d, 0 <= t < rhi, against a box.  NaN-safe in the conservative direction:
// fminf/fmaxf drop NaN operands (0 * inf on a slab boundary), and a ray whose direction is NaN (dead slot) is
// filtered by the caller's live mask.
__device__ __forceinline__ bool ray_reaches_box(float ox, float oy, float oz, float dx, float dy, float dz, float rhi, const float4& lo, const float4& hi) {
    const float ix = rcp_approx(dx), iy = rcp_approx(dy), iz = rcp_approx(dz);
    const float x1 = (lo.x - ox) * ix, x2 = (hi.x - ox) * ix;
    const float y1 = (lo.y - oy) * iy, y2 = (hi.y - oy) * iy;
    const float z1 = (lo.z - oz) * iz, z2 = (hi.z - oz) * iz;
    const float tmin = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fmaxf(fminf(z1, z2), 0.0f));
    const float tmax = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fminf(fmaxf(z1, z2), rhi));
    // relative slack for the approximate reciprocals / products (the box already carries an absolute margin)
    return !(tmin > tmax * 1.0001f + 1e-30f);
}

// ------------------------------------------------------------------------------------------------
// Exact re-evaluation (cold path)
// ------------------------------------------------------------------------------------------------
// xyz = intersection point, w = Vec3Df::distance(origin, I) (raytracing.cpp:182) or -1 for "no hit".
__device__ __noinline__ float4 exact_eval_tri(const float4* __restrict__ triv, int tri, float ox, float oy, float oz, float tx, float ty, float tz) {
    const float4 A = __ldg(&triv[3 * tri]), B = __ldg(&triv[3 * tri + 1]), C = __ldg(&triv[3 * tri + 2]);
    v3 I;
    const v3 O = mk3(ox, oy, oz);
    if (!exact_ray_triangle(O, mk3(tx, ty, tz), mk3(A), mk3(B), mk3(C), I)) return make_float4(0.f, 0.f, 0.f, -1.0f);
    return make_float4(I.x, I.y, I.z, e_distance(O, I));
}

// ------------------------------------------------------------------------------------------------
// Tile pipeline: triangle records stream through a kStages-deep shared-memory ring filled by 1-D TMA bulk
// copies.  There is no producer warp: the LAST warp to finish a stage re-arms its mbarrier and issues the
// copy for the tile kStages iterations ahead, so nobody ever blocks on an "empty" barrier.
// ------------------------------------------------------------------------------------------------
template <int VEC> struct __align__(128) ScanSmemV {
    float4 tiles[kStages][kTile * VEC];
    unsigned long long full[kStages];
    unsigned int done[kStages];
};
using ScanSmem = ScanSmemV<kRecVec>;
// tile culling only (CULL = true kernels): tiles some ray of this CTA can reach (bitmap), and the batch being streamed
struct CullSmem {
    unsigned int need[kCullMaxTiles / 32];
    unsigned short list[kCullList];
    unsigned int list_n, list_next;
};


// Per-ray state that only the cold path and the epilogues touch (nearest distance, nearest triangle, sample id).
// In the brute-force kernels it lives in shared memory, one column per thread, which frees 3*R registers for the
// filter loop; the culling kernels (whose shared memory is taken by the tile list) keep it in registers.
template <int R> struct ColdSmem { float dist[R][kThreads]; int best[R][kThreads]; uint32_t sid[R][kThreads]; };
template <class T> struct SmemCol {
    T (*col)[kThreads];
    __device__ __forceinline__ T& operator[](int k) const { return col[k][threadIdx.x]; }
};
template <class T, int R> struct RegCol {
    T v[R];
    __device__ __forceinline__ T& operator[](int k) { return v[k]; }
    __device__ __forceinline__ const T& operator[](int k) const { return v[k]; }
};
template <int R, bool IN_SMEM> struct ColdState;
template <int R> struct ColdState<R, true> {
    SmemCol<float> dist; SmemCol<int> best; SmemCol<uint32_t> sid;
    __device__ __forceinline__ explicit ColdState(ColdSmem<R>* s) : dist{s->dist}, best{s->best}, sid{s->sid} {}
};
template <int R> struct ColdState<R, false> {
    RegCol<float, R> dist; RegCol<int, R> best; RegCol<uint32_t, R> sid;
    __device__ __forceinline__ explicit ColdState(ColdSmem<R>*) {}
};
template <int R, bool PRESENT> struct ColdStorage { ColdSmem<R> s; __device__ ColdSmem<R>* get() { return &s; } };
template <int R> struct ColdStorage<R, false> { __device__ ColdSmem<R>* get() { return nullptr; } };

// Shared-memory block of the culling path, present only in the CULL = true kernel instantiations.
template <bool CULL> struct CullStorage { CullSmem s; __device__ CullSmem& get() { return s; } __device__ const unsigned short* list() { return s.list; } };
template <> struct CullStorage<false> { __device__ CullSmem& get() { return *reinterpret_cast<CullSmem*>(this); } __device__ const unsigned short* list() { return nullptr; } };

// Everything a scan kernel keeps in shared memory, as ONE block of dynamic shared memory (the 8-rays-per-thread pencil
// shape needs 57 KB: more than the 48 KB a kernel may declare statically).
template <int R, bool CULL, int VEC = kRecVec> struct KernelSmem {
    ScanSmemV<VEC> sm;
    CullStorage<CULL> csm;
    ColdStorage<R, !CULL> cold;
};
extern __shared__ __align__(128) unsigned char rt_dyn_smem[];

// Work decomposition of one scan launch.  count rays -> nchunks chunks of kThreads*R rays; when there are fewer chunks
// than CTAs the triangle tiles are split into `parts` ranges of `len` tiles each (the record array carries kPadTiles
// "never" tiles, so the last range may run past ntiles).  Work item w = chunk * parts + part; CTA b takes items
// b, b + gridDim.x, ...  Results of the parts of a chunk are merged through atomics on per-ray keys / lit bits.
struct Split {
    uint32_t nchunks, parts, len, nitems;
};
__device__ __forceinline__ Split make_split(uint32_t count, uint32_t per_chunk, int ntiles, bool allow_split) {
    Split sp;
    sp.nchunks = (count + per_chunk - 1) / per_chunk;
    uint32_t parts = 1;
    if (allow_split && sp.nchunks > 0 && sp.nchunks < kItemsPerCta * gridDim.x) {
        // aim at >= kItemsPerCta items per CTA so that the last (partial) round of the static item striding costs
        // a few per cent at most, but keep >= kMinPartTiles tiles per part to amortise the per-item ray set-up;
        // a launch with fewer chunks than CTAs is always split as far as the tiles allow
        parts = (kItemsPerCta * gridDim.x + sp.nchunks - 1) / sp.nchunks;
        const uint32_t cap = (sp.nchunks < gridDim.x) ? (uint32_t)ntiles : ((uint32_t)ntiles + kMinPartTiles - 1) / kMinPartTiles;
        if (parts > cap) parts = cap;
        if (parts > (uint32_t)kMaxParts) parts = kMaxParts;
        if (parts < 1) parts = 1;
    }
    sp.len = ((uint32_t)ntiles + parts - 1) / parts;
    sp.parts = ((uint32_t)ntiles + sp.len - 1) / sp.len;   // no empty part
    sp.nitems = sp.nchunks * sp.parts;
    return sp;
}

struct Pipe {
    uint32_t tiles_addr, full_addr;
    unsigned int* done;
    const float4* tiles_ptr;
    const float4* rec;
    uint32_t vec;          // float4 per record of the streamed array (kRecVec, or kTpVec for the thread-pencil records)
    uint32_t parts, len;   // see Split
    uint32_t total_iters;  // tiles this CTA will consume over its whole life (range mode) / end of the current batch (list mode)
    uint32_t it;           // next iteration
    const unsigned short* list;  // list mode (hierarchical culling): tile of iteration i is list[i - base]; nullptr = range mode
    uint32_t base;
};

template <bool LIST>
__device__ __forceinline__ void pipe_issue(const Pipe& p, uint32_t iter) {
    const uint32_t stage = iter % kStages;
    const uint32_t bar = p.full_addr + stage * 8u;
    uint32_t tile;
    if (LIST) {
        tile = p.list[iter - p.base];
    } else {
        const uint32_t q = iter / p.len;                               // this CTA's q-th work item
        const uint32_t item = blockIdx.x + q * gridDim.x;
        tile = (item % p.parts) * p.len + (iter - q * p.len);
    }
    // generic-proxy reads of this stage (all warps are past it) are ordered before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t bytes = kTile * p.vec * (uint32_t)sizeof(float4);
    mbar_arrive_expect_tx(bar, bytes);
    tma_bulk_g2s(p.tiles_addr + stage * bytes, p.rec + (size_t)tile * kTile * p.vec, bytes, bar);
}

__device__ __forceinline__ uint32_t cta_items(uint32_t nitems) {
    return (nitems > blockIdx.x) ? (nitems - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
}

// list_mode: nothing is prefetched here; every batch is started by pipe_begin_batch().
template <bool LIST, int VEC = kRecVec>
__device__ __forceinline__ void pipe_init(Pipe& p, ScanSmemV<VEC>& sm, const float4* rec, const Split& sp, const unsigned short* list) {
    p.vec = VEC;
    p.tiles_addr = smem_u32(&sm.tiles[0][0]);
    p.full_addr = smem_u32(&sm.full[0]);
    p.done = sm.done;
    p.tiles_ptr = &sm.tiles[0][0];
    p.rec = rec;
    p.parts = sp.parts;
    p.len = sp.len;
    p.total_iters = LIST ? 0u : cta_items(sp.nitems) * sp.len;
    p.it = 0;
    p.list = list;
    p.base = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(p.full_addr + s * 8u, 1); sm.done[s] = 0; }
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t n = p.total_iters < (uint32_t)kStages ? p.total_iters : (uint32_t)kStages;
        for (uint32_t i = 0; i < n; ++i) pipe_issue<LIST>(p, i);
    }
}

// List mode: the n tiles of the list are the next n iterations.  Call from all threads after a __syncthreads() that
// follows both the list construction and every warp's release of the previous batch.
__device__ __forceinline__ void pipe_begin_batch(Pipe& p, uint32_t n) {
    p.base = p.it;
    p.total_iters = p.it + n;
    if (threadIdx.x == 0) {
        const uint32_t m = n < (uint32_t)kStages ? n : (uint32_t)kStages;
        for (uint32_t i = 0; i < m; ++i) pipe_issue<true>(p, p.it + i);
    }
}

__device__ __forceinline__ const float4* pipe_acquire(const Pipe& p) {
    const uint32_t stage = p.it % kStages;
    mbar_wait(p.full_addr + stage * 8u, (p.it / kStages) & 1u);
    return p.tiles_ptr + stage * (kTile * p.vec);
}

template <bool LIST>
__device__ __forceinline__ void pipe_release(Pipe& p) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t stage = p.it % kStages;
        const unsigned int prev = atomicAdd(&p.done[stage], 1u);
        if ((prev % kWarps) == kWarps - 1) {  // every warp has finished reading this stage
            const uint32_t nxt = p.it + kStages;
            if (nxt < p.total_iters) pipe_issue<LIST>(p, nxt);
        }
    }
    ++p.it;
}

// ------------------------------------------------------------------------------------------------
// The scan: R = 2*RP rays per thread against every triangle.
// ------------------------------------------------------------------------------------------------
template <int RP>
struct FastRays {
    float2 ox[RP], oy[RP], oz[RP];  // shifted origins O' = O - eps_r * d
    float2 dx[RP], dy[RP], dz[RP];  // unit directions d
    uint32_t rhi[2 * RP];           // bits of (best distance + 2*eps_r); 0 = ray finished / unused
};

template <int RP, int K>
__device__ __forceinline__ void fast_set(FastRays<RP>& f, v3 O, v3 D, float eps_r, bool live) {
    // d = normalize(D - O) only has to be accurate to a few ulp: it feeds the filter, never a result.
    float dx = D.x - O.x, dy = D.y - O.y, dz = D.z - O.z;
    const float len2 = dx * dx + dy * dy + dz * dz;
    float inv = rsqrtf(len2);
    // a (near-)zero or overflowing direction cannot be normalised: d = 0 makes cos = 0 for every triangle, i.e. every
    // pair goes to the exact path (|cos| < bmin), which is always sound
    if (!(len2 > 1e-30f) || !(len2 < 1e30f)) inv = 0.0f;
    dx *= inv; dy *= inv; dz *= inv;
    // an unused / finished slot gets a NaN direction: cos = NaN fails every clause of the candidate test
    if (!live) { dx = dy = dz = __int_as_float(0x7fc00000); O = mk3(0.f, 0.f, 0.f); }
    constexpr int p = K / 2;
    if (K & 1) {
        f.dx[p].y = dx; f.dy[p].y = dy; f.dz[p].y = dz;
        f.ox[p].y = O.x - eps_r * dx; f.oy[p].y = O.y - eps_r * dy; f.oz[p].y = O.z - eps_r * dz;
    } else {
        f.dx[p].x = dx; f.dy[p].x = dy; f.dz[p].x = dz;
        f.ox[p].x = O.x - eps_r * dx; f.oy[p].x = O.y - eps_r * dy; f.oz[p].x = O.z - eps_r * dz;
    }
    f.rhi[K] = live ? 0x7f7fffffu : 0u;  // FLT_MAX: the reference starts from dist = FLT_MAX (raytracing.cpp:164)
}

// One filter evaluation for a pair of rays against one record of dominant-axis class W; returns the two candidate
// predicates.  16 packed FP32 instructions + 2 MUFU.RCP + the compare logic.
// GRAZ = false: the scene-level proof of rt_b200.cu (no_grazing) says that every pair with |cos| < cos_min is rejected
// by the reference itself (its |b| < 1e-5 test, raytracing.cpp:115), so the grazing clause is compiled out.
template <int RP, int W, bool GRAZ>
__device__ __forceinline__ void filter_pair(const FastRays<RP>& f, int p, const float4& q0, const float4& q1, const float4& q2,
                                            const float4& q3, bool& c0, bool& c1) {
    const float2 ou = (W == 0) ? f.oy[p] : (W == 1) ? f.oz[p] : f.ox[p];
    const float2 ov = (W == 0) ? f.oz[p] : (W == 1) ? f.ox[p] : f.oy[p];
    const float2 du = (W == 0) ? f.dy[p] : (W == 1) ? f.dz[p] : f.dx[p];
    const float2 dv = (W == 0) ? f.dz[p] : (W == 1) ? f.dx[p] : f.dy[p];
    float2 b = __fmul2_rn(splat2(q0.x), f.dx[p]);
    b = __ffma2_rn(splat2(q0.y), f.dy[p], b);
    b = __ffma2_rn(splat2(q0.z), f.dz[p], b);
    float2 a = __ffma2_rn(splat2(q0.x), f.ox[p], splat2(q0.w));
    a = __ffma2_rn(splat2(q0.y), f.oy[p], a);
    a = __ffma2_rn(splat2(q0.z), f.oz[p], a);
    const float2 rc = make_float2(rcp_approx(-b.x), rcp_approx(-b.y));
    const float2 r = __fmul2_rn(a, rc);
    const float2 iu = __ffma2_rn(r, du, ou);
    const float2 iv = __ffma2_rn(r, dv, ov);
    float2 s = __ffma2_rn(splat2(q1.x), iu, splat2(q1.z));
    s = __ffma2_rn(splat2(q1.y), iv, s);
    float2 t = __ffma2_rn(splat2(q2.x), iu, splat2(q2.z));
    t = __ffma2_rn(splat2(q2.y), iv, t);
    float2 q = __fadd2_rn(splat2(q1.w), make_float2(-s.x, -s.y));
    q = __fadd2_rn(q, make_float2(-t.x, -t.y));
    const float m0 = fminf(fminf(s.x, t.x), q.x), m1 = fminf(fminf(s.y, t.y), q.y);
    const float2 e = __fmul2_rn(splat2(q2.w), rc);  // tolerance -E1*|1/cos| = -|e|
    // a NaN in s/t/q/e (non-finite geometry) keeps the pair a candidate; a NaN cos (dead ray slot) never is one
    c0 = !(m0 < -fabsf(e.x)) && (__float_as_uint(r.x) < f.rhi[2 * p]);
    c1 = !(m1 < -fabsf(e.y)) && (__float_as_uint(r.y) < f.rhi[2 * p + 1]);
    if (GRAZ) {
        c0 = c0 || (fabsf(b.x) < q3.x);
        c1 = c1 || (fabsf(b.y) < q3.x);
    }
}

// Finished ray (shadow ray that found its occluder): no further candidates.
template <int RP>
__device__ __forceinline__ void kill_slot(FastRays<RP>& f, int k) {
    const float nan = __int_as_float(0x7fc00000);
    if (k & 1) { f.dx[k / 2].y = nan; f.dy[k / 2].y = nan; f.dz[k / 2].y = nan; }
    else { f.dx[k / 2].x = nan; f.dy[k / 2].x = nan; f.dz[k / 2].x = nan; }
    f.rhi[k] = 0u;
}

// ------------------------------------------------------------------------------------------------
// Pencil filter (rt_pencil.h): rays through a common point, in the launch's projective chart.  6 registers per ray
// pair; the hot loop evaluates the three weights only -- 6 packed FP32 instructions and 3 LOP3 per (ray pair,
// triangle) -- and ANDs the sign words per ray ("any candidate in the block" = some live ray's AND has its sign bit
// clear); the cold path rebuilds the block's mask with the full test (+ 3 packed instructions for the distance clause).
// ------------------------------------------------------------------------------------------------
template <int RP>
struct PencilRays {
    float2 x[RP], y[RP];    // chart coordinates of the direction away from the common point: dir / (dir.f) = (x, y, 1); NaN = not representable (every triangle exact)
    float2 zhi[RP];         // depth bound of the distance clause: depth of the origin + (nearest distance so far + s_lam) / |(x, y, 1)|
    uint32_t dead[2 * RP];  // 0x80000000: finished / unused slot (its sign words are ignored)
};
__device__ __forceinline__ float& pair_elem(float2& v, int odd) { return odd ? v.y : v.x; }

__device__ __forceinline__ float pencil_depth(const FrameParams& P, v3 O) {   // (O - E).f
    return fmaf(O.x - P.pE[0], P.pF[6], fmaf(O.y - P.pE[1], P.pF[7], (O.z - P.pE[2]) * P.pF[8]));
}
// depth bound for a ray whose nearest accepted distance so far is `nearest` (0 for an any-hit shadow ray: r >= 0)
__device__ __forceinline__ float pencil_zhi(const FrameParams& P, v3 O, float x, float y, float nearest) {
    const float c = __fmul_ru(rsqrtf(fmaf(x, x, fmaf(y, y, 1.0f))), 1.000002f);   // >= 1/|(x, y, 1)|: depth gained per unit of distance
    float z = __fadd_ru(pencil_depth(P, O), __fmul_ru(__fadd_ru(nearest, P.p_lam_slack), c));
    if (!(z < FLT_MAX)) z = FLT_MAX;   // "nothing yet" (nearest = FLT_MAX) and NaN (ray outside the chart): no bound
    return z;
}

// flip: the pencil's centre is the ray's DEST (shadow rays), the direction away from it is origin - dest.
// Returns false when the ray cannot be represented in the chart (does not point into the chart's half space, too far
// off axis, not finite); the slot then holds NaN coordinates = candidate for every triangle.
template <int RP>
__device__ __forceinline__ bool pencil_set_slot(PencilRays<RP>& f, int k, const FrameParams& P, v3 O, v3 D, bool flip, float nearest0, bool live) {
    const float s = flip ? -1.0f : 1.0f;
    const float dx = s * (D.x - O.x), dy = s * (D.y - O.y), dz = s * (D.z - O.z);   // the sign of a float difference is exact
    float x, y;
    const bool ok = pencil_chart_xy(P.pF, P.p_wmax2, dx, dy, dz, x, y);
    if (!ok) x = y = __int_as_float(0x7fc00000);
    float z = pencil_zhi(P, O, x, y, nearest0);
    if (!live) { x = y = 0.0f; z = 0.0f; }
#pragma unroll
    for (int kk = 0; kk < 2 * RP; ++kk)
        if (kk == k) {
            pair_elem(f.x[kk / 2], kk & 1) = x; pair_elem(f.y[kk / 2], kk & 1) = y; pair_elem(f.zhi[kk / 2], kk & 1) = z;
            f.dead[kk] = live ? 0u : 0x80000000u;
        }
    return ok;
}

// sign words of one ray pair against one pencil record: bit 31 of s0 / s1 set <=> certainly not a candidate
template <int RP>
__device__ __forceinline__ void pencil_pair(const PencilRays<RP>& f, int p, const float4& q0, const float4& q1, const float4& q2, const float4& q3,
                                            uint32_t& s0, uint32_t& s1) {
    float2 a = __ffma2_rn(splat2(q0.y), f.y[p], splat2(q0.z));
    float2 b = __ffma2_rn(splat2(q1.y), f.y[p], splat2(q1.z));
    float2 c = __ffma2_rn(splat2(q2.y), f.y[p], splat2(q2.z));
    a = __ffma2_rn(splat2(q0.x), f.x[p], a);
    b = __ffma2_rn(splat2(q1.x), f.x[p], b);
    c = __ffma2_rn(splat2(q2.x), f.x[p], c);
    float2 sg = __ffma2_rn(splat2(q1.w), f.y[p], splat2(q2.w));
    sg = __ffma2_rn(splat2(q0.w), f.x[p], sg);
    const float2 e = __ffma2_rn(sg, f.zhi[p], splat2(q3.x));
    s0 = __float_as_uint(a.x) | __float_as_uint(b.x) | __float_as_uint(c.x) | __float_as_uint(e.x);
    s1 = __float_as_uint(a.y) | __float_as_uint(b.y) | __float_as_uint(c.y) | __float_as_uint(e.y);
}

// Hot-loop form: the three weights only (6 packed FP32 instructions + 1 LOP3 per ray), no distance clause.  The pairs it
// lets through that the full test would stop (triangles pierced by the ray's line behind the nearest hit / beyond a shadow
// ray's origin) cost a cold-path entry, not an exact evaluation: the cold path rebuilds its mask with pencil_pair().
template <int RP>
__device__ __forceinline__ void pencil_pair_hot(const PencilRays<RP>& f, int p, const float4& q0, const float4& q1, const float4& q2, uint32_t& s0, uint32_t& s1) {
    float2 a = __ffma2_rn(splat2(q0.y), f.y[p], splat2(q0.z));
    float2 b = __ffma2_rn(splat2(q1.y), f.y[p], splat2(q1.z));
    float2 c = __ffma2_rn(splat2(q2.y), f.y[p], splat2(q2.z));
    a = __ffma2_rn(splat2(q0.x), f.x[p], a);
    b = __ffma2_rn(splat2(q1.x), f.x[p], b);
    c = __ffma2_rn(splat2(q2.x), f.x[p], c);
    s0 = __float_as_uint(a.x) | __float_as_uint(b.x) | __float_as_uint(c.x);
    s1 = __float_as_uint(a.y) | __float_as_uint(b.y) | __float_as_uint(c.y);
}

// The same three weights for ONE ray with scalar FMAs (3 register operands per instruction instead of 4-5): with 8 rays per
// thread the record loads amortise and the loop is issue-bound rather than register-bandwidth-bound (tools/ubench/pencil.cu:
// 4.21e12 against 3.59e12 tests/s for the packed 4-ray shape).
__device__ __forceinline__ uint32_t pencil_ray_hot(float x, float y, const float4& q0, const float4& q1, const float4& q2) {
    const float a = fmaf(q0.x, x, fmaf(q0.y, y, q0.z));
    const float b = fmaf(q1.x, x, fmaf(q1.y, y, q1.z));
    const float c = fmaf(q2.x, x, fmaf(q2.y, y, q2.z));
    return __float_as_uint(a) | __float_as_uint(b) | __float_as_uint(c);
}

// Ray-side hooks of the cold path, one overload per filter.
struct GenericBand { float eps_r2; };   // what the ray-side hook of the generic filter needs
template <int RP>
__device__ __forceinline__ void ray_nearer(FastRays<RP>& f, int k, float dist, const GenericBand& band, v3) { f.rhi[k] = __float_as_uint(__fadd_ru(dist, band.eps_r2)); }
template <int RP>
__device__ __forceinline__ void ray_nearer(PencilRays<RP>& f, int k, float dist, const FrameParams& P, v3 O) {
#pragma unroll
    for (int kk = 0; kk < 2 * RP; ++kk)
        if (kk == k) {
            float& zhi = pair_elem(f.zhi[kk / 2], kk & 1);
            const float v = pencil_zhi(P, O, pair_elem(f.x[kk / 2], kk & 1), pair_elem(f.y[kk / 2], kk & 1), dist);
            if (v < zhi) zhi = v;   // a ray outside the chart (NaN) keeps "no bound"
        }
}
template <int RP>
__device__ __forceinline__ void ray_kill(FastRays<RP>& f, int k) { kill_slot<RP>(f, k); }
template <int RP>
__device__ __forceinline__ void ray_kill(PencilRays<RP>& f, int k) {
#pragma unroll
    for (int kk = 0; kk < 2 * RP; ++kk)
        if (kk == k) f.dead[kk] = 0x80000000u;
}

template <int RP, int J>
struct BitLayout {
    static constexpr int R = 2 * RP;
    static_assert(R * J <= 32, "candidate mask must fit 32 bits");
    // bit j*R set for every j < J:  (2^(R*J) - 1) / (2^R - 1)
    static constexpr uint32_t kRep = (uint32_t)((((uint64_t)1 << (R * J)) - 1) / (((uint64_t)1 << R) - 1));
};

// Cold path of one filter block: exact re-evaluation of every (ray, triangle) pair in `mask` (bit j*R + k).
// NEAREST: keeps (dist, best) exactly like intersectMesh; !NEAREST: any-hit, clears the ray's live bit on the first
// exact hit.  fetch(k, O, D) returns the exact ray of slot k.
template <int RP, int J, bool NEAREST, class Rays, class Band, class DistT, class BestT, class Fetch>
__device__ __forceinline__ void exact_block(uint32_t mask, const float4* rec, int jb, Rays& fr, DistT& dist, BestT& best, uint32_t& live,
                                            const float4* __restrict__ triv, const Band& band, const Fetch& fetch, uint32_t& n_exact) {
    constexpr int R = 2 * RP;
    constexpr uint32_t REP = BitLayout<RP, J>::kRep;
    mask &= live * REP;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const uint32_t mk = (mask >> k) & REP;
        if (mk) {
            v3 O, D;
            fetch(k, O, D);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if ((mk & (1u << (j * R))) && (NEAREST || ((live >> k) & 1u))) {
                    const int tri = (int)__float_as_uint(rec[(jb + j) * kRecVec + 3].y);   // triangle id of this record
                    const float4 e = exact_eval_tri(triv, tri, O.x, O.y, O.z, D.x, D.y, D.z);
                    ++n_exact;
                    if (NEAREST) {
                        // intersectMesh keeps the first (= lowest-index) triangle among equal distances (strict <,
                        // raytracing.cpp:183); records are not visited in index order, so the index breaks ties here
                        if (!(e.w < 0.0f) && (e.w < dist[k] || (e.w == dist[k] && tri < best[k]))) {
                            dist[k] = e.w;
                            best[k] = tri;
                            ray_nearer<RP>(fr, k, e.w, band, O);
                        }
                    } else {
                        // isShadow ends with index != -1 iff some hit had distance < FLT_MAX
                        // (raytracing.cpp:164,183); a NaN / inf distance never registers
                        if (e.w >= 0.0f && e.w < FLT_MAX && (live & (1u << k))) {
                            live &= ~(1u << k);
                            best[k] = tri;
                            ray_kill<RP>(fr, k);
                        }
                    }
                }
            }
        }
    }
}

// One tile (kTile records of dominant-axis class W) against this thread's R rays, generic filter.
template <int RP, int J, bool NEAREST, int W, bool GRAZ, class DistT, class BestT, class Fetch>
__device__ __forceinline__ void scan_tile(const float4* rec, FastRays<RP>& fr, DistT& dist, BestT& best, uint32_t& live,
                                          const float4* __restrict__ triv, float eps_r2, const Fetch& fetch, uint32_t& n_exact) {
    constexpr int R = 2 * RP;
    const int nvalid = __float_as_int(rec[3].z);   // records in use in this tile (warp-uniform); the rest is padding
#pragma unroll 1
    for (int jb = 0; jb < nvalid; jb += J) {
        // hot: only "is there any candidate in this block" (the predicate ORs fold into the compares)
        bool any = false;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 q0 = rec[(jb + j) * kRecVec + 0], q1 = rec[(jb + j) * kRecVec + 1];
            const float4 q2 = rec[(jb + j) * kRecVec + 2], q3 = rec[(jb + j) * kRecVec + 3];
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                bool c0, c1;
                filter_pair<RP, W, GRAZ>(fr, p, q0, q1, q2, q3, c0, c1);
                any = any || c0 || c1;
            }
        }
        if (any) {  // cold: redo the block's filter to find which pairs, then exact re-evaluation per ray
            uint32_t mask = 0;
#pragma unroll 1
            for (int j = 0; j < J; ++j) {
                const float4 q0 = rec[(jb + j) * kRecVec + 0], q1 = rec[(jb + j) * kRecVec + 1];
                const float4 q2 = rec[(jb + j) * kRecVec + 2], q3 = rec[(jb + j) * kRecVec + 3];
#pragma unroll
                for (int p = 0; p < RP; ++p) {
                    bool c0, c1;
                    filter_pair<RP, W, GRAZ>(fr, p, q0, q1, q2, q3, c0, c1);
                    mask |= ((c0 ? 1u : 0u) << (2 * p) | (c1 ? 1u : 0u) << (2 * p + 1)) << (j * R);
                }
            }
            exact_block<RP, J, NEAREST>(mask, rec, jb, fr, dist, best, live, triv, GenericBand{eps_r2}, fetch, n_exact);
        }
    }
}

// The same against a tile of PENCIL records (no axis classes).
template <int RP, int J, bool NEAREST, class DistT, class BestT, class Fetch>
__device__ __forceinline__ void scan_tile_pencil(const float4* rec, PencilRays<RP>& fr, DistT& dist, BestT& best, uint32_t& live,
                                                 const float4* __restrict__ triv, const FrameParams& P, const Fetch& fetch, uint32_t& n_exact) {
    constexpr int R = 2 * RP;
    const int nvalid = __float_as_int(rec[3].z);
#pragma unroll 1
    for (int jb = 0; jb < nvalid; jb += J) {
        uint32_t acc[R];   // per ray: AND of the sign words; bit 31 survives iff every triangle of the block is rejected
#pragma unroll
        for (int k = 0; k < R; ++k) acc[k] = 0xffffffffu;
        static_assert(J % 2 == 0, "the hot loop folds two triangles per AND");
#pragma unroll
        for (int j = 0; j < J; j += 2) {   // two triangles per step: acc & u & u' is one LOP3
            const float4 q0 = rec[(jb + j) * kRecVec + 0], q1 = rec[(jb + j) * kRecVec + 1], q2 = rec[(jb + j) * kRecVec + 2];
            const float4 r0 = rec[(jb + j + 1) * kRecVec + 0], r1 = rec[(jb + j + 1) * kRecVec + 1], r2 = rec[(jb + j + 1) * kRecVec + 2];
            if constexpr (RP >= 3) {   // scalar form (same IEEE FMAs, same order, same bits as the packed one)
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const float x = (k & 1) ? fr.x[k / 2].y : fr.x[k / 2].x, y = (k & 1) ? fr.y[k / 2].y : fr.y[k / 2].x;
                    acc[k] &= pencil_ray_hot(x, y, q0, q1, q2) & pencil_ray_hot(x, y, r0, r1, r2);
                }
            } else {
#pragma unroll
                for (int p = 0; p < RP; ++p) {
                    uint32_t s0, s1, u0, u1;
                    pencil_pair_hot<RP>(fr, p, q0, q1, q2, s0, s1);
                    pencil_pair_hot<RP>(fr, p, r0, r1, r2, u0, u1);
                    acc[2 * p] &= s0 & u0;
                    acc[2 * p + 1] &= s1 & u1;
                }
            }
        }
        uint32_t all = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < R; ++k) all &= acc[k] | fr.dead[k];
        if ((int)all >= 0) {
            uint32_t mask = 0;
#pragma unroll 1
            for (int j = 0; j < J; ++j) {
                const float4 q0 = rec[(jb + j) * kRecVec + 0], q1 = rec[(jb + j) * kRecVec + 1];
                const float4 q2 = rec[(jb + j) * kRecVec + 2], q3 = rec[(jb + j) * kRecVec + 3];
#pragma unroll
                for (int p = 0; p < RP; ++p) {
                    uint32_t s0, s1;
                    pencil_pair<RP>(fr, p, q0, q1, q2, q3, s0, s1);
                    mask |= ((~s0 >> 31) << (2 * p) | (~s1 >> 31) << (2 * p + 1)) << (j * R);
                }
            }
            exact_block<RP, J, NEAREST>(mask, rec, jb, fr, dist, best, live, triv, P, fetch, n_exact);
        }
    }
}

// Does any live ray of this thread reach the box?
template <int RP>
__device__ __forceinline__ bool thread_reaches_box(const FastRays<RP>& fr, uint32_t live, const float4& lo, const float4& hi) {
    bool need = false;
#pragma unroll
    for (int k = 0; k < 2 * RP; ++k) {
        const int p = k / 2;
        const bool odd = k & 1;
        const bool reach = ray_reaches_box(odd ? fr.ox[p].y : fr.ox[p].x, odd ? fr.oy[p].y : fr.oy[p].x, odd ? fr.oz[p].y : fr.oz[p].x,
                                           odd ? fr.dx[p].y : fr.dx[p].x, odd ? fr.dy[p].y : fr.dy[p].x, odd ? fr.dz[p].y : fr.dz[p].x,
                                           __uint_as_float(fr.rhi[k]), lo, hi);
        need = need || (reach && ((live >> k) & 1u));
    }
    return need;
}

// Tile culling (RT_OPT_TILE_CULLING, scenes of <= kCullMaxTiles tiles; larger scenes are scanned brute force) for one work item:
//   1. every warp walks the super-tile boxes (kSuper tiles each) of the item's tile range and, inside the super-tiles
//      one of its rays reaches, the tile boxes; reached tiles are marked in a CTA-wide bitmap;
//   2. the bitmap is compacted into batches of <= kCullList tile ids;
//   3. only those tiles are streamed through the TMA ring; a warp still re-tests a tile's box against its own rays
//      (with the nearest-hit bound as it is by then) before scanning it.
// Same filter + exact tiers as scan_pass on every tile that is not skipped, so the results are identical.
template <int RP, int J, bool NEAREST, bool GRAZ, class DistT, class BestT, class Fetch>
__device__ __forceinline__ void scan_item_culled(Pipe& pipe, CullSmem& sm, FastRays<RP>& fr, DistT& dist, BestT& best, uint32_t& live,
                                                 const float4* __restrict__ triv, float eps_r2, const Fetch& fetch, uint32_t& n_exact, int tile_begin,
                                                 int tile_end, const float4* __restrict__ tile_box, const float4* __restrict__ super_box, int cls1, int cls2) {
    const int lane = threadIdx.x & 31;
    const int w0 = tile_begin / 32, w1 = (tile_end + 31) / 32;   // bitmap words of the range
    __syncthreads();   // every warp is done with the previous item's bitmap / list / ring
    for (int w = w0 + (int)threadIdx.x; w < w1; w += kThreads) sm.need[w] = 0u;
    if (threadIdx.x == 0) sm.list_next = (unsigned int)w0;
    __syncthreads();
    // 1. mark
    for (int g = tile_begin / kSuper; g * kSuper < tile_end; ++g) {
        const float4 slo = __ldg(&super_box[2 * g]), shi = __ldg(&super_box[2 * g + 1]);
        if (!__any_sync(0xffffffffu, thread_reaches_box<RP>(fr, live, slo, shi))) continue;
        const int t0 = max(g * kSuper, tile_begin), t1 = min((g + 1) * kSuper, tile_end);
        unsigned int bits = 0u;
        for (int t = t0; t < t1; ++t) {
            const float4 lo = __ldg(&tile_box[2 * t]), hi = __ldg(&tile_box[2 * t + 1]);
            if (__any_sync(0xffffffffu, thread_reaches_box<RP>(fr, live, lo, hi))) bits |= 1u << (t & 31);
        }
        if (lane == 0 && bits) atomicOr(&sm.need[g * kSuper / 32], bits);   // kSuper == 32: one word per super-tile
    }
    // 2. + 3. batches
    for (;;) {
        __syncthreads();   // marks complete (first round) / previous batch fully consumed and released
        if (threadIdx.x < 32) {   // warp 0 compacts the next batch, word by word
            unsigned int n = 0;
            int w = (int)sm.list_next;
            while (w < w1 && n + 32 <= (unsigned int)kCullList) {
                unsigned int word = sm.need[w];
                const bool mine = (word >> lane) & 1u;
                const unsigned int pos = n + __popc(word & ((1u << lane) - 1u));
                if (mine) sm.list[pos] = (unsigned short)(w * 32 + lane);
                n += __popc(word);
                ++w;
            }
            if (lane == 0) { sm.list_n = n; sm.list_next = (unsigned int)w; }
        }
        __syncthreads();
        const uint32_t n = sm.list_n;
        if (n == 0 && (int)sm.list_next >= w1) break;
        pipe_begin_batch(pipe, n);
        for (uint32_t i = 0; i < n; ++i) {
            const float4* rec = pipe_acquire(pipe);
            const int tile = (int)sm.list[i];
            const float4 lo = __ldg(&tile_box[2 * tile]), hi = __ldg(&tile_box[2 * tile + 1]);
            if (__any_sync(0xffffffffu, thread_reaches_box<RP>(fr, live, lo, hi))) {
                if (tile < cls1) scan_tile<RP, J, NEAREST, 0, GRAZ>(rec, fr, dist, best, live, triv, eps_r2, fetch, n_exact);
                else if (tile < cls2) scan_tile<RP, J, NEAREST, 1, GRAZ>(rec, fr, dist, best, live, triv, eps_r2, fetch, n_exact);
                else scan_tile<RP, J, NEAREST, 2, GRAZ>(rec, fr, dist, best, live, triv, eps_r2, fetch, n_exact);
            }
            pipe_release<true>(pipe);
        }
        if ((int)sm.list_next >= w1) break;
    }
}

// Scans the tiles [tile_begin, tile_begin + pipe.len) of one work item (brute force: every tile).  NEAREST: keeps
// (dist, best) like intersectMesh; !NEAREST: any-hit, a ray dies at its first exact hit.  fetch(k, O, D) returns the
// exact ray of slot k.  cls1 / cls2: first tile of class 1 / class 2 (tiles are grouped by the dominant axis of their triangles).
template <int RP, int J, bool NEAREST, bool GRAZ, class DistT, class BestT, class Fetch>
__device__ __forceinline__ void scan_pass(Pipe& pipe, FastRays<RP>& fr, DistT& dist, BestT& best, uint32_t& live,
                                          const float4* __restrict__ triv, float eps_r2, const Fetch& fetch, uint32_t& n_exact, int tile_begin,
                                          int cls1, int cls2) {
    for (int tile = tile_begin; tile < tile_begin + (int)pipe.len; ++tile) {
        const float4* rec = pipe_acquire(pipe);
        // warp-level early exit: shadow rays that all found their occluder
        if (__any_sync(0xffffffffu, live != 0u)) {
            if (tile < cls1) scan_tile<RP, J, NEAREST, 0, GRAZ>(rec, fr, dist, best, live, triv, eps_r2, fetch, n_exact);
            else if (tile < cls2) scan_tile<RP, J, NEAREST, 1, GRAZ>(rec, fr, dist, best, live, triv, eps_r2, fetch, n_exact);
            else scan_tile<RP, J, NEAREST, 2, GRAZ>(rec, fr, dist, best, live, triv, eps_r2, fetch, n_exact);
        }
        pipe_release<false>(pipe);
    }
}

template <int RP, int J, bool NEAREST, class DistT, class BestT, class Fetch>
__device__ __forceinline__ void scan_pass_pencil(Pipe& pipe, PencilRays<RP>& fr, DistT& dist, BestT& best, uint32_t& live,
                                                 const float4* __restrict__ triv, const FrameParams& P, const Fetch& fetch, uint32_t& n_exact, int tile_begin) {
    for (int tile = tile_begin; tile < tile_begin + (int)pipe.len; ++tile) {
        const float4* rec = pipe_acquire(pipe);
        if (__any_sync(0xffffffffu, live != 0u)) scan_tile_pencil<RP, J, NEAREST>(rec, fr, dist, best, live, triv, P, fetch, n_exact);
        pipe_release<false>(pipe);
    }
}

// ------------------------------------------------------------------------------------------------
// Ray generation, main.cpp:380-386 (exact)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ v3 lerp_corner(const float* c, int off, float xs, float omx, float ys, float omy) {
    // yscale*(xscale*c00 + (1-xscale)*c10) + (1-yscale)*(xscale*c01 + (1-xscale)*c11)
    const v3 c00 = mk3(c[0 + off], c[1 + off], c[2 + off]), c01 = mk3(c[6 + off], c[7 + off], c[8 + off]);
    const v3 c10 = mk3(c[12 + off], c[13 + off], c[14 + off]), c11 = mk3(c[18 + off], c[19 + off], c[20 + off]);
    const v3 top = e_scale(e_add(e_scale(c00, xs), e_scale(c10, omx)), ys);
    const v3 bot = e_scale(e_add(e_scale(c01, xs), e_scale(c11, omx)), omy);
    return e_add(top, bot);
}

// Primary work order.  Slot i -> sample: pixels are visited in 8x8 blocks (Morton order inside a block, blocks row-major),
// the sub-samples of a pixel innermost, so that every warp / CTA / chunk of consecutive slots -- and, through the
// compaction order, of secondary rays -- covers a compact screen region (coherent rays make the shadow early-exit and
// tile culling effective).  Slots of a partial block that fall outside the chunk's pixels are invalid.
__device__ __forceinline__ bool slot_to_sample(const FrameParams& P, uint32_t slot, uint32_t& s) {
    if (P.trace_api) { s = slot; return slot < P.nsamples; }
    const uint32_t spp = P.pfx * P.pfy;
    const uint32_t t = slot / spp, sub = slot - t * spp;
    const uint32_t blk = t >> 6, m = t & 63u;
    const uint32_t by = blk / P.tiles_x, bx = blk - by * P.tiles_x;
    // Morton decode of 6 bits: x = bits 0,2,4   y = bits 1,3,5
    const uint32_t mx = (m & 1u) | ((m >> 1) & 2u) | ((m >> 2) & 4u);
    const uint32_t my = ((m >> 1) & 1u) | ((m >> 2) & 2u) | ((m >> 3) & 4u);
    const uint32_t x = bx * 8u + mx, ry = by * 8u + my;
    s = (ry * P.W + x) * spp + sub;
    return x < P.W && ry < P.nrows;
}

__device__ __forceinline__ void primary_ray(const FrameParams& P, uint32_t s, v3& O, v3& D) {
    const uint32_t spp = P.pfx * P.pfy;
    const uint32_t pix = s / spp, sub = s - pix * spp;
    const uint32_t subx = sub / P.pfy, suby = sub - subx * P.pfy;
    const uint32_t ry = pix / P.W, x = pix - ry * P.W;
    const uint32_t y = (P.row0 + ry) * P.G + P.rank;
    const float xs = __fsub_rn(1.0f, __fdiv_rn(__fadd_rn(__fmul_rn((float)x, (float)P.pfx), (float)(int)subx), P.divX));  // main.cpp:380
    const float ys = __fsub_rn(1.0f, __fdiv_rn(__fadd_rn(__fmul_rn((float)y, (float)P.pfy), (float)(int)suby), P.divY));  // main.cpp:381
    const float omx = __fsub_rn(1.0f, xs), omy = __fsub_rn(1.0f, ys);
    O = lerp_corner(P.corners, 0, xs, omx, ys, omy);
    D = lerp_corner(P.corners, 3, xs, omx, ys, omy);
}

// ------------------------------------------------------------------------------------------------
// k_trace: nearest hit for primary rays (generated in registers) or queued continuation rays.
// ------------------------------------------------------------------------------------------------
// Ray set-up shared by the scan kernels: slot k of this thread gets ray (O, D) or is marked unused.
template <int RP>
__device__ __forceinline__ void fast_set_slot(FastRays<RP>& fr, int k, v3 O, v3 D, float eps_r, bool ok) {
    constexpr int R = 2 * RP;
    // (fast_set needs a compile-time slot)
    if (k == 0) fast_set<RP, 0>(fr, O, D, eps_r, ok);
    if (k == 1) fast_set<RP, 1>(fr, O, D, eps_r, ok);
    if (k == 2) fast_set<RP, (R > 2 ? 2 : 0)>(fr, O, D, eps_r, ok);
    if (k == 3) fast_set<RP, (R > 2 ? 3 : 0)>(fr, O, D, eps_r, ok);
    if (k == 4) fast_set<RP, (R > 4 ? 4 : 0)>(fr, O, D, eps_r, ok);
    if (k == 5) fast_set<RP, (R > 4 ? 5 : 0)>(fr, O, D, eps_r, ok);
    if (k == 6) fast_set<RP, (R > 6 ? 6 : 0)>(fr, O, D, eps_r, ok);
    if (k == 7) fast_set<RP, (R > 6 ? 7 : 0)>(fr, O, D, eps_r, ok);
}

template <bool B, class X, class Y> struct SelectT { using type = X; };
template <class X, class Y> struct SelectT<false, X, Y> { using type = Y; };

// PENCIL (brute force): the pencil filter around the eye for the primary rays of a frame (P.prec, P.pE), or around the mirror
// image of the eye for the queue of a plane group's level-1 continuation rays (P.mirror_sel >= 0; k_shade checked every one).
template <int RP, int J, int MINB, bool PRIMARY, bool GRAZ, bool CULL, bool PENCIL = false>
__global__ void __launch_bounds__(kThreads, MINB) k_trace(const __grid_constant__ FrameParams P, int level) {
    static_assert(!PENCIL || !CULL, "the pencil filter serves the brute-force scan");
    constexpr int R = 2 * RP;
    KernelSmem<R, CULL>& ks = *reinterpret_cast<KernelSmem<R, CULL>*>(rt_dyn_smem);
    ScanSmem& sm = ks.sm;
    CullStorage<CULL>& csm = ks.csm;
    ColdStorage<R, !CULL>& cold = ks.cold;
    const uint32_t count = PRIMARY ? P.nslots : ray_queue_count(P, level);
    const uint32_t* __restrict__ queue = ray_queue(P);
    const uint32_t per_chunk = kThreads * R;
    const Split sp = make_split(count, per_chunk, P.ntiles, true);
    Pipe pipe;
    pipe_init<CULL>(pipe, sm, PENCIL ? P.prec : P.rec, sp, csm.list());
    uint32_t n_exact = 0;
    const float eps_r2 = 2.0f * P.eps_r;

    for (uint32_t item = blockIdx.x; item < sp.nitems; item += gridDim.x) {
        const uint32_t chunk = item / sp.parts, part = item - chunk * sp.parts;
        typename SelectT<PENCIL, PencilRays<RP>, FastRays<RP>>::type fr;
        ColdState<R, !CULL> st(cold.get());
        auto& dist = st.dist; auto& best = st.best; auto& sid = st.sid;
        uint32_t live = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            // a thread's R rays are consecutive slots (neighbouring sub-samples / pixels)
            const uint32_t ray = chunk * per_chunk + threadIdx.x * R + k;
            bool ok = ray < count;
            v3 O = mk3(0, 0, 0), D = mk3(0, 0, 1);
            uint32_t s = 0;
            if (PRIMARY && ok) ok = slot_to_sample(P, ray, s);
            if (ok) {
                if (PRIMARY && !P.trace_api) {
                    primary_ray(P, s, O, D);
                    if (part == 0) {  // every part regenerates the ray (cheap); one of them publishes it
                        P.ray_o[s] = make_float4(O.x, O.y, O.z, 0.f);
                        P.ray_d[s] = make_float4(D.x, D.y, D.z, __int_as_float(0));
                        P.thr[s] = make_float4(1.f, 1.f, 1.f, 0.f);
                        P.acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                } else {
                    if (!PRIMARY) s = queue[ray];
                    const float4 o = P.ray_o[s], d = P.ray_d[s];
                    O = mk3(o); D = mk3(d);
                }
                live |= 1u << k;
            }
            sid[k] = s;
            dist[k] = FLT_MAX;
            best[k] = -1;
            if constexpr (PENCIL) pencil_set_slot<RP>(fr, k, P, O, D, false, FLT_MAX, ok);   // outside the chart: every triangle exact
            else fast_set_slot<RP>(fr, k, O, D, P.eps_r, ok);
        }
        // fetch() re-reads a ray for the exact path: primary rays of other parts may not be published yet
        auto fetch = [&](int k, v3& O, v3& D) {
            if (PRIMARY && !P.trace_api) { primary_ray(P, sid[k], O, D); }
            else { const float4 o = P.ray_o[sid[k]], d = P.ray_d[sid[k]]; O = mk3(o); D = mk3(d); }
        };
        if constexpr (PENCIL)
            scan_pass_pencil<RP, J, true>(pipe, fr, dist, best, live, P.triv, P, fetch, n_exact, (int)(part * sp.len));
        else if constexpr (CULL)
            scan_item_culled<RP, J, true, GRAZ>(pipe, csm.get(), fr, dist, best, live, P.triv, eps_r2, fetch, n_exact, (int)(part * sp.len),
                                                min((int)((part + 1) * sp.len), P.ntiles), P.tile_box, P.super_box, P.cls1, P.cls2);
        else
            scan_pass<RP, J, true, GRAZ>(pipe, fr, dist, best, live, P.triv, eps_r2, fetch, n_exact, (int)(part * sp.len), P.cls1, P.cls2);

        // merge: (distance bits, triangle id) -- the smallest distance wins, equal distances go to the lowest
        // index, which is exactly the sequential rule of intersectMesh (strict <, raytracing.cpp:183)
#pragma unroll
        for (int k = 0; k < R; ++k)
            if (((live >> k) & 1u) && best[k] >= 0)
                atomicMin(&P.key[sid[k]], ((unsigned long long)__float_as_uint(dist[k]) << 32) | (unsigned int)best[k]);
    }
    if (n_exact) atomicAdd(reinterpret_cast<unsigned long long*>(&P.counters[kCntExact]), (unsigned long long)n_exact);
}

// ------------------------------------------------------------------------------------------------
// k_finish: one thread per ray of this level.  Reads the merged nearest-triangle key, recomputes the exact hit
// point of the winner, tests the analytic spheres (after the triangles, strict <), writes the hit record,
// arms the lit bits for k_shadow and appends hits to the queue (warp-aggregated).
// ------------------------------------------------------------------------------------------------
// (tid, stride): this thread's index among the `stride` threads that share the work -- the whole grid for k_finish, one
// CTA for the single-launch path of small rt_trace batches (k_trace_small); every thread of a warp must call it.
template <bool PRIMARY>
__device__ __forceinline__ void finish_rays(const FrameParams& P, int level, uint32_t tid, uint32_t stride) {
    const uint32_t count = PRIMARY ? P.nslots : ray_queue_count(P, level);
    const uint32_t* __restrict__ queue = ray_queue(P);
    const uint32_t rounds = (count + stride - 1) / stride;
    const uint32_t all_lit = (P.nlights >= 32) ? 0xffffffffu : ((1u << P.nlights) - 1u);
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t i = r * stride + tid;
        bool hit_any = false;
        uint32_t s = 0;
        bool ok = i < count;
        if (ok) { if (PRIMARY) ok = slot_to_sample(P, i, s); else s = queue[i]; }
        if (ok) {
            const unsigned long long key = P.key[s];
            P.key[s] = kKeyEmpty;  // ready for the next level / frame
            int idx = (key == kKeyEmpty) ? -1 : (int)(unsigned int)(key & 0xffffffffull);
            float dbest = (key == kKeyEmpty) ? FLT_MAX : __uint_as_float((unsigned int)(key >> 32));
            const float4 o = P.ray_o[s], d = P.ray_d[s];
            const v3 O = mk3(o), D = mk3(d);
            v3 I = mk3(0, 0, 0);
            if (idx >= 0) {
                const float4 e = exact_eval_tri(P.triv, idx, O.x, O.y, O.z, D.x, D.y, D.z);
                I = mk3(e);
            }
            for (int a = 0; a < P.n_always; ++a) {   // triangles outside the filter: same (distance, id) rule as the scan
                const int tri = (int)P.always_list[a];
                const float4 e = exact_eval_tri(P.triv, tri, O.x, O.y, O.z, D.x, D.y, D.z);
                if (!(e.w < 0.0f) && (e.w < dbest || (e.w == dbest && tri < idx))) { dbest = e.w; idx = tri; I = mk3(e); }
            }
            for (int sp = 0; sp < P.nspheres; ++sp) {  // spheres come after the triangles, strict <
                const float4 c = P.spheres[2 * sp];
                v3 Is;
                if (exact_ray_sphere(O, D, mk3(c), c.w, Is)) {
                    const float ds = e_distance(O, Is);
                    if (ds < dbest) { dbest = ds; idx = P.ntri + sp; I = Is; }
                }
            }
            P.hit[s] = make_float4(I.x, I.y, I.z, __int_as_float(idx));
            P.lit[s] = all_lit;    // k_shadow clears the bit of every light that is occluded
            if (PRIMARY && P.prim_out) P.prim_out[P.sample_base + s] = idx;
            hit_any = idx >= 0;
        }
        warp_append(hit_any, s, P.q_hit, &P.counters[kCntHit + level]);
        if (P.n_always > 0) {
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if ((threadIdx.x & 31) == 0 && m) atomicAdd(reinterpret_cast<unsigned long long*>(&P.counters[kCntExact]), (unsigned long long)__popc(m) * (unsigned long long)P.n_always);
        }
    }
}
template <bool PRIMARY>
__global__ void __launch_bounds__(256) k_finish(const __grid_constant__ FrameParams P, int level) {
    finish_rays<PRIMARY>(P, level, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ------------------------------------------------------------------------------------------------
// k_shadow: one work item = (hit sample, light).  isShadow, raytracing.cpp:241-261.
//   ANY     : exact iff no material has (has_Tr && Tr < 1): any occluder shadows, warp-level early exit;
//             the triangle range may be split over CTAs (an occluder found by any part clears the lit bit).
//   NEAREST : the nearest occluder's material decides (transparent -> lit); never split.
// lit[] arrives with every light's bit set (k_finish); occluded lights are cleared here.
// ------------------------------------------------------------------------------------------------
// P.light_sel >= 0: the launch serves that light only (one launch per light: pencil launches, and the generic launches of
// the lights that do not qualify in a frame that uses the pencil filter).
// PENCIL (any-hit, brute force): the pencil filter around the light.  Directions run from the light to the hit point,
// occluders between the two have 0 < depth <= depth of the origin; occluders BEYOND the light (the reference's shadow
// rays are unbounded) cannot exist for a ray inside the chart's half space; a ray outside the chart skips the scan and is
// tested exactly against every triangle in the epilogue (rare by construction: pencil_light_setup only accepts lights
// outside the scene box, and origins inside the box + bias are always inside the chart).
template <int RP, int J, int MINB, bool NEAREST, bool GRAZ, bool CULL, bool PENCIL = false>
__global__ void __launch_bounds__(kThreads, MINB) k_shadow(const __grid_constant__ FrameParams P, int level) {
    static_assert(!PENCIL || (!NEAREST && !CULL), "the pencil filter serves any-hit shadow rays in the brute-force scan");
    constexpr int R = 2 * RP;
    KernelSmem<R, CULL>& ks = *reinterpret_cast<KernelSmem<R, CULL>*>(rt_dyn_smem);
    ScanSmem& sm = ks.sm;
    CullStorage<CULL>& csm = ks.csm;
    ColdStorage<R, !CULL>& cold = ks.cold;
    const int lsel = P.light_sel;
    const uint32_t nl = lsel >= 0 ? 1u : (uint32_t)P.nlights;
    const uint32_t count = P.counters[kCntHit + level] * nl;
    const uint32_t per_chunk = kThreads * R;
    const Split sp = make_split(count, per_chunk, P.ntiles, !NEAREST);
    Pipe pipe;
    pipe_init<CULL>(pipe, sm, PENCIL ? P.prec : P.rec, sp, csm.list());
    uint32_t n_exact = 0;
    const float eps_r2 = 2.0f * P.eps_r;

    for (uint32_t item = blockIdx.x; item < sp.nitems; item += gridDim.x) {
        const uint32_t chunk = item / sp.parts, part = item - chunk * sp.parts;
        typename SelectT<PENCIL, PencilRays<RP>, FastRays<RP>>::type fr;
        ColdState<R, !CULL> st(cold.get());
        auto& dist = st.dist; auto& best = st.best; auto& sid = st.sid;
        uint32_t live = 0, valid = 0, unsafe = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const uint32_t ray = chunk * per_chunk + threadIdx.x * R + k;
            const bool ok = ray < count;
            bool scan = ok;
            v3 O = mk3(0, 0, 0), D = mk3(0, 0, 1);
            sid[k] = 0;
            if (ok) {
                const uint32_t h = ray / nl, l = lsel >= 0 ? (uint32_t)lsel : ray - h * nl;
                sid[k] = P.q_hit[h];
                O = e_add(mk3(P.hit[sid[k]]), mk3(0.1f, 0.1f, 0.1f));
                D = mk3(P.lights[l][0], P.lights[l][1], P.lights[l][2]);
                valid |= 1u << k;
            }
            dist[k] = FLT_MAX;
            best[k] = -1;
            if constexpr (PENCIL) {
                // a ray outside the chart (it leaves the light's half space, so an occluder BEYOND the light is possible, or
                // it is too far off axis) is not scanned: the epilogue tests it exactly against every triangle
                if (!pencil_set_slot<RP>(fr, k, P, O, D, true, 0.0f, ok) && ok) { unsafe |= 1u << k; scan = false; fr.dead[k] = 0x80000000u; }
            } else {
                fast_set_slot<RP>(fr, k, O, D, P.eps_r, ok);
            }
            if (scan) live |= 1u << k;
        }
        const uint32_t ray0 = chunk * per_chunk + threadIdx.x * R;   // slot k of this thread is work item ray0 + k
        auto fetch = [&](int k, v3& O, v3& D) {   // the exact shadow ray of slot k: hit + (0.1, 0.1, 0.1) -> light (raytracing.cpp:246-248)
            const uint32_t l = lsel >= 0 ? (uint32_t)lsel : (ray0 + k) % nl;
            O = e_add(mk3(P.hit[sid[k]]), mk3(0.1f, 0.1f, 0.1f));
            D = mk3(P.lights[l][0], P.lights[l][1], P.lights[l][2]);
        };
        if constexpr (PENCIL)
            scan_pass_pencil<RP, J, false>(pipe, fr, dist, best, live, P.triv, P, fetch, n_exact, (int)(part * sp.len));
        else if constexpr (CULL)
            scan_item_culled<RP, J, NEAREST, GRAZ>(pipe, csm.get(), fr, dist, best, live, P.triv, eps_r2, fetch, n_exact, (int)(part * sp.len),
                                                   min((int)((part + 1) * sp.len), P.ntiles), P.tile_box, P.super_box, P.cls1, P.cls2);
        else
            scan_pass<RP, J, NEAREST, GRAZ>(pipe, fr, dist, best, live, P.triv, eps_r2, fetch, n_exact, (int)(part * sp.len), P.cls1, P.cls2);

#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (!((valid >> k) & 1u)) continue;
            bool lit;
            v3 O, D;
            fetch(k, O, D);
            if (NEAREST) {
                int idx = best[k];
                float dbest = dist[k];
                for (int a = 0; a < P.n_always; ++a) {
                    const int tri = (int)P.always_list[a];
                    const float4 e = exact_eval_tri(P.triv, tri, O.x, O.y, O.z, D.x, D.y, D.z);
                    if (!(e.w < 0.0f) && (e.w < dbest || (e.w == dbest && tri < idx))) { dbest = e.w; idx = tri; }
                }
                for (int sph = 0; sph < P.nspheres; ++sph) {
                    const float4 c = P.spheres[2 * sph];
                    v3 Is;
                    if (exact_ray_sphere(O, D, mk3(c), c.w, Is)) {
                        const float ds = e_distance(O, Is);
                        if (ds < dbest) { dbest = ds; idx = P.ntri + sph; }
                    }
                }
                if (idx < 0) {
                    lit = true;
                } else {
                    const uint32_t m = (idx < P.ntri) ? __float_as_uint(P.normal_mat[idx].w) : __float_as_uint(P.spheres[2 * (idx - P.ntri) + 1].x);
                    const float4 ks_tr = P.materials[4 * m + 2];
                    const uint32_t flags = __float_as_uint(P.materials[4 * m + 3].x);
                    lit = (flags & RT_HAS_TR) && (ks_tr.w < 1.0f);  // transparent occluder: no shadow (:254)
                }
            } else {
                lit = (live >> k) & 1u;  // still alive after every triangle of this part: no occluder found here
                if (PENCIL && ((unsafe >> k) & 1u)) {   // not scanned: every triangle exactly, once (part 0)
                    lit = true;
                    if (part == 0)
                        for (int tri = 0; lit && tri < P.ntri; ++tri) {
                            const float4 e = exact_eval_tri(P.triv, tri, O.x, O.y, O.z, D.x, D.y, D.z);
                            if (e.w >= 0.0f && e.w < FLT_MAX) lit = false;
                        }
                    if (part == 0) n_exact += (uint32_t)P.ntri;
                }
                if (part == 0) {
                    for (int a = 0; lit && a < P.n_always; ++a) {
                        const float4 e = exact_eval_tri(P.triv, (int)P.always_list[a], O.x, O.y, O.z, D.x, D.y, D.z);
                        if (e.w >= 0.0f && e.w < FLT_MAX) lit = false;
                    }
                    for (int sph = 0; lit && sph < P.nspheres; ++sph) {
                        const float4 c = P.spheres[2 * sph];
                        v3 Is;
                        if (exact_ray_sphere(O, D, mk3(c), c.w, Is) && e_distance(O, Is) < FLT_MAX) lit = false;
                    }
                }
            }
            if (!lit) atomicAnd(&P.lit[sid[k]], ~(1u << (lsel >= 0 ? (uint32_t)lsel : (ray0 + k) % nl)));
        }
    }
    if (n_exact) atomicAdd(reinterpret_cast<unsigned long long*>(&P.counters[kCntExact]), (unsigned long long)n_exact);
}

// ------------------------------------------------------------------------------------------------
// k_shade: shade() for every hit of this level, then the continuation ray (reflection XOR refraction).
// The colour recursion of the reference is a linear chain c0 + K0*(c1 + K1*(...)); it is evaluated
// iteratively with a throughput thr = K0*K1*..., acc += thr*c (float rounding differs from the nested
// form by ~1e-7 relative, far inside the 1/255 tolerance; geometry -- every ray origin/dest -- is exact).
// ------------------------------------------------------------------------------------------------
struct Mat {
    v3 Kd, Ka, Ks;
    float Ns, Ni, Tr;
    uint32_t flags;
};
__device__ __forceinline__ Mat load_material(const float4* __restrict__ mats, uint32_t m) {
    const float4 a = mats[4 * m], b = mats[4 * m + 1], c = mats[4 * m + 2], d = mats[4 * m + 3];
    Mat M;
    M.Kd = mk3(a); M.Ns = a.w; M.Ka = mk3(b); M.Ni = b.w; M.Ks = mk3(c); M.Tr = c.w; M.flags = __float_as_uint(d.x);
    return M;
}

// reflection() + addOffset(), raytracing.cpp:277-285, 266-271. `ray` is normalised here (again).
__device__ __forceinline__ void reflect_ray(v3 ray, v3 Ppos, v3 normal, v3& point, v3& dest) {
    ray = e_normalize(ray);
    const v3 Rv = e_sub(ray, e_scale(normal, __fmul_rn(2.0f, e_dot(normal, ray))));
    dest = e_add(Ppos, Rv);
    v3 off = e_normalize(e_sub(dest, Ppos));
    off = e_scale(off, 0.01f);
    point = e_add(Ppos, off);
}
__device__ __forceinline__ void offset_point(v3 Ppos, v3 dest, v3& point) {  // addOffset
    v3 off = e_normalize(e_sub(dest, Ppos));
    off = e_scale(off, 0.01f);
    point = e_add(Ppos, off);
}

__device__ __forceinline__ void shade_hits(const FrameParams& P, int level, uint32_t tid, uint32_t stride) {
    const uint32_t count = P.counters[kCntHit + level];
    const uint32_t rounds = (count + stride - 1) / stride;
    const bool fAmbient = P.features & RT_AMBIENT, fDiffuse = P.features & RT_DIFFUSE, fSpecular = P.features & RT_SPECULAR;
    const bool fReflection = P.features & RT_REFLECTION, fShadows = P.features & RT_SHADOWS, fRefraction = P.features & RT_REFRACTION;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t h = r * stride + tid;
        bool spawn = false;
        int group = -1;   // >= 0: the continuation ray goes to the queue of that plane group's mirror pencil
        uint32_t s = 0;
        if (h < count) {
            s = P.q_hit[h];
            const float4 ro = P.ray_o[s], rd = P.ray_d[s], hp = P.hit[s];
            const int lvl = __float_as_int(rd.w);
            const int idx = __float_as_int(hp.w);
            const v3 Ppos = mk3(hp);
            const v3 ray = e_sub(mk3(rd), mk3(ro));  // raytracing.cpp:393
            v3 normal;
            uint32_t mi;
            if (idx < P.ntri) {
                const float4 nm = P.normal_mat[idx];
                normal = mk3(nm);                     // never flipped toward the ray (:394)
                mi = __float_as_uint(nm.w);
            } else {
                const float4 c = P.spheres[2 * (idx - P.ntri)];
                normal = e_normalize(e_sub(Ppos, mk3(c)));  // Sphere.h:33-37
                mi = __float_as_uint(P.spheres[2 * (idx - P.ntri) + 1].x);
            }
            const Mat M = load_material(P.materials, mi);
            const v3 cam = mk3(P.camera[0], P.camera[1], P.camera[2]);

            v3 c = mk3(0.f, 0.f, 0.f);
            if (fAmbient && (M.flags & RT_HAS_KA)) c = e_add(c, M.Ka);                       // :337-340
            const uint32_t litbits = fShadows ? P.lit[s] : 0xffffffffu;
            for (int i = 0; i < P.nlights; ++i) {
                if (!((litbits >> i) & 1u)) continue;                                          // isShadow :345
                const v3 L = mk3(P.lights[i][0], P.lights[i][1], P.lights[i][2]);
                if (fDiffuse && (M.flags & RT_HAS_KD)) {                                       // diffuseOnly :197-205
                    normal = e_normalize(normal);
                    const v3 lp = e_normalize(L);  // the light POSITION as a direction
                    v3 d = e_add(mk3(0.f, 0.f, 0.f), e_scale(M.Kd, std_max(e_dot(normal, lp), 0.0f)));
                    c = e_add(c, e_scale(d, M.Tr));
                }
                if (fSpecular && (M.flags & RT_HAS_KS) && (M.flags & RT_HAS_NS)) {             // blinnPhong :210-232
                    v3 V = e_sub(cam, Ppos);
                    normal = e_normalize(normal);
                    V = e_normalize(V);
                    v3 Lv = e_normalize(e_sub(L, Ppos));
                    v3 H = e_normalize(e_add(V, Lv));
                    float spec = std_max(e_dot(H, normal), 0.0f);
                    spec = powf(spec, M.Ns);       // CUDA powf vs glibc powf: a few ulp, colour only
                    v3 sp = e_add(mk3(0.f, 0.f, 0.f), e_scale(M.Ks, spec));
                    c = e_add(c, e_scale(sp, M.Tr));
                }
            }
            const float4 th = P.thr[s];
            float4 ac = P.acc[s];
            ac.x = __fadd_rn(ac.x, __fmul_rn(th.x, c.x));
            ac.y = __fadd_rn(ac.y, __fmul_rn(th.y, c.y));
            ac.z = __fadd_rn(ac.z, __fmul_rn(th.z, c.z));
            P.acc[s] = ac;

            // continuation: refraction XOR reflection (:357-364)
            v3 point = Ppos, dest = Ppos, K = mk3(0.f, 0.f, 0.f);
            int nlvl = lvl;
            if (fRefraction && (M.Tr < 1.0f) && lvl < P.max_lvl) {                            // refraction(..., lvl+1) :290-330
                const int L1 = lvl + 1;
                const v3 rn = e_normalize(ray);
                const float check = e_dot(rn, normal);
                if (check < 0.0f) {
                    // :297-298  `angle = acosf(check); if (angle <= 2 && angle > 0)`.  CUDA's acosf is not glibc's, and a 1-ulp
                    // difference at angle == 2 would flip reflect <-> refract for that sample.  The branch is decided without
                    // the libm call: for check < 0, glibc's acosf(check) <= 2.0f  <=>  check >= -0.416146934f (bits 0xbed51136),
                    // checked exhaustively over every negative float in [-1, 0) against glibc 2.39 (acosf is monotone there;
                    // tests/test_oracle_golden.py repeats the check around the threshold); angle > 0 always holds for
                    // check < 0, and check < -1 (rounding) gives NaN <= 2 = false on the CPU, check >= T = false here.
                    if (check >= __uint_as_float(0xbed51136u)) {                                // :298 grazing hack
                        reflect_ray(rn, Ppos, normal, point, dest);
                        K = M.Ks; nlvl = L1 + 1; spawn = true;
                    } else {
                        const float nr = __fdiv_rn(1.0f, M.Ni);
                        const float dn = e_dot(normal, rn);
                        float root = __fsub_rn(1.0f, __fmul_rn(__fmul_rn(nr, nr), __fsub_rn(1.0f, __fmul_rn(dn, dn))));
                        if (root >= 0.0f) {
                            root = __fsqrt_rn(root);
                            const v3 T = e_sub(e_scale(e_sub(rn, e_scale(normal, dn)), nr), e_scale(normal, root));  // :307
                            dest = e_add(Ppos, T);
                            offset_point(Ppos, dest, point);
                            const float k = __fsub_rn(1.0f, M.Tr);
                            K = mk3(k, k, k); nlvl = L1 + 1; spawn = true;
                        }
                    }
                } else {
                    const float nr = M.Ni;
                    const v3 nn = e_neg(normal);
                    const float dn = e_dot(nn, rn);
                    float root = __fsub_rn(1.0f, __fmul_rn(__fmul_rn(nr, nr), __fsub_rn(1.0f, __fmul_rn(dn, dn))));
                    if (root >= 0.0f) {
                        root = __fsqrt_rn(root);
                        const v3 T = e_sub(e_scale(e_sub(rn, e_scale(nn, dn)), nr), e_scale(nn, root));                // :321
                        dest = e_add(Ppos, T);
                        offset_point(Ppos, dest, point);
                        const float k = __fsub_rn(1.0f, M.Tr);
                        K = mk3(k, k, k); nlvl = L1 + 1; spawn = true;
                    }
                }
            } else if (fReflection && lvl < P.max_lvl) {                                      // Ks * reflection(..., lvl+1)
                reflect_ray(ray, Ppos, normal, point, dest);
                K = M.Ks; nlvl = lvl + 1; spawn = true;   // traced even when Ks == 0, like the reference
                // reflection of a PRIMARY ray off a triangle of a plane group: the mirror pencil of that plane takes the ray if
                // (checked here, on the ray as built) its line passes through the mirror image of the eye (rt_pencil.h)
                if ((P.n_mirrors > 0 || P.tp_on) && level == 0 && idx < P.ntri) {
                    const float Of[3] = {point.x, point.y, point.z}, Df[3] = {dest.x, dest.y, dest.z};
                    if (P.n_mirrors > 0) {
                        const uint32_t g = P.tri_group[idx];
                        if (g < (uint32_t)P.n_mirrors && pencil_mirror_accepts(P.mirror[g], Of, Df)) group = (int)g;
                    }
                    // ... or off any other triangle: its own mirror image of the eye, shared by the rays of a thread (rt_tpencil.h)
                    if (group < 0 && P.tp_on) {
                        const float4 A = P.triv[3 * idx], B = P.triv[3 * idx + 1], C = P.triv[3 * idx + 2];
                        const float Af[3] = {A.x, A.y, A.z}, Bf[3] = {B.x, B.y, B.z}, Cf[3] = {C.x, C.y, C.z};
                        float E[3];
                        TpRay tr;
                        if (tp_mirror_point(P.tp.eye, P.tp.centerf, Af, Bf, Cf, E) && tp_accepts(P.tp, E, Of, Df, tr)) {
                            group = kQueueTp;
                            atomicAdd(&P.tp_hist[idx], 1u);
                        }
                    }
                }
            }
            if (spawn) {
                P.ray_o[s] = make_float4(point.x, point.y, point.z, 0.f);
                P.ray_d[s] = make_float4(dest.x, dest.y, dest.z, __int_as_float(nlvl));
                P.thr[s] = make_float4(__fmul_rn(th.x, K.x), __fmul_rn(th.y, K.y), __fmul_rn(th.z, K.z), 0.f);
            }
        }
        warp_append(spawn && group < 0, s, P.q_ray, &P.counters[kCntRay + level + 1]);
        if (P.tp_on && level == 0) warp_append(spawn && group == kQueueTp, s, P.tp_pool, &P.counters[kCntTpPool]);
        for (int g = 0; g < P.n_mirrors; ++g)   // (0 unless this is level 0 of a frame with mirror pencils)
            warp_append(spawn && group == g, s, P.q_mirror + (size_t)g * P.q_mirror_stride, &P.counters[kCntMirror + g]);
    }
}
__global__ void __launch_bounds__(256) k_shade(const __grid_constant__ FrameParams P, int level) {
    shade_hits(P, level, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ------------------------------------------------------------------------------------------------
// k_trace_small: the whole wavefront of a SMALL rt_trace batch in ONE launch of ONE CTA -- the drop-in
// performRayTracing(origin, dest) call (raytracing.cpp:410-416) is a batch of one ray, and a dozen-level recursion costs
// 4 launches per level otherwise.  No filter: with a handful of rays every (ray, triangle) pair simply goes through
// exact_eval_tri, the triangles split over the CTA's threads, (distance, id) keys merged exactly like the scan kernels
// merge theirs; finish_rays / shade_hits are the bodies of k_finish / k_shade, run by the same CTA between barriers
// (global memory written by the CTA is visible to it after __syncthreads()).  hit0: level-0 hit records, copied aside.
// ------------------------------------------------------------------------------------------------
constexpr int kSmallThreads = 512;   // 128 registers per thread: shade_hits does not spill
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
__global__ void __launch_bounds__(kSmallThreads, 1) k_trace_small(const __grid_constant__ FrameParams P, int levels, float4* __restrict__ hit0, int shadow_nearest) {
    __shared__ unsigned long long s_key;       // shadow ray in flight: (distance, id) of its nearest occluder
    const uint32_t tid = threadIdx.x, T = blockDim.x;
    const bool shadows = (P.features & RT_SHADOWS) && P.nlights > 0;
    for (int level = 0; level < levels; ++level) {
        const uint32_t count = level == 0 ? P.nslots : P.counters[kCntRay + level];
        if (count == 0) break;   // uniform: nothing left to trace
        // (1) nearest hit of every queued ray: triangles over threads
        for (uint32_t r = 0; r < count; ++r) {
            const uint32_t s = level == 0 ? r : P.q_ray[r];
            const float4 o = P.ray_o[s], d = P.ray_d[s];
            unsigned long long best = kKeyEmpty;
            for (int tri = (int)tid; tri < P.ntri; tri += (int)T) {
                const float4 e = exact_eval_tri(P.triv, tri, o.x, o.y, o.z, d.x, d.y, d.z);
                if (e.w >= 0.0f && e.w < FLT_MAX) {   // what registers in intersectMesh (strict < from FLT_MAX, raytracing.cpp:164,183)
                    const unsigned long long k = ((unsigned long long)__float_as_uint(e.w) << 32) | (unsigned int)tri;
                    best = k < best ? k : best;
                }
            }
            best = warp_min_u64(best);
            if ((tid & 31u) == 0 && best != kKeyEmpty) atomicMin(&P.key[s], best);
        }
        if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&P.counters[kCntExact]), (unsigned long long)count * (unsigned long long)P.ntri);
        __syncthreads();
        // (2) hit records (spheres, always-exact list: none outstanding -- every triangle was evaluated exactly; the list is re-evaluated harmlessly)
        if (level == 0) finish_rays<true>(P, 0, tid, T); else finish_rays<false>(P, level, tid, T);
        __syncthreads();
        const uint32_t nhit = P.counters[kCntHit + level];
        if (level == 0 && hit0)
            for (uint32_t i = tid; i < P.nsamples; i += T) hit0[i] = P.hit[i];
        // (3) shadow rays: isShadow (raytracing.cpp:241-261), one (hit, light) pair at a time
        if (shadows) {
            for (uint32_t h = 0; h < nhit; ++h) {
                const uint32_t s = P.q_hit[h];
                const v3 O = e_add(mk3(P.hit[s]), mk3(0.1f, 0.1f, 0.1f));
                for (int l = 0; l < P.nlights; ++l) {
                    const v3 D = mk3(P.lights[l][0], P.lights[l][1], P.lights[l][2]);
                    if (tid == 0) s_key = kKeyEmpty;
                    __syncthreads();
                    unsigned long long best = kKeyEmpty;
                    for (int tri = (int)tid; tri < P.ntri; tri += (int)T) {
                        const float4 e = exact_eval_tri(P.triv, tri, O.x, O.y, O.z, D.x, D.y, D.z);
                        if (e.w >= 0.0f && e.w < FLT_MAX) {
                            const unsigned long long k = ((unsigned long long)__float_as_uint(e.w) << 32) | (unsigned int)tri;
                            best = k < best ? k : best;
                        }
                    }
                    for (int sp = (int)tid; sp < P.nspheres; sp += (int)T) {   // spheres come after the triangles, strict <: id = ntri + sp loses ties
                        const float4 c = P.spheres[2 * sp];
                        v3 Is;
                        if (exact_ray_sphere(O, D, mk3(c), c.w, Is)) {
                            const float ds = e_distance(O, Is);
                            if (ds >= 0.0f && ds < FLT_MAX) {
                                const unsigned long long k = ((unsigned long long)__float_as_uint(ds) << 32) | (unsigned int)(P.ntri + sp);
                                best = k < best ? k : best;
                            }
                        }
                    }
                    best = warp_min_u64(best);
                    if ((tid & 31u) == 0 && best != kKeyEmpty) atomicMin(&s_key, best);
                    __syncthreads();
                    if (tid == 0) {
                        const unsigned long long k = s_key;
                        bool lit = true;
                        if (k != kKeyEmpty) {
                            lit = false;
                            if (shadow_nearest) {   // the nearest occluder's material decides: transparent -> no shadow (:253-256)
                                const int idx = (int)(unsigned int)(k & 0xffffffffull);
                                const uint32_t m = (idx < P.ntri) ? __float_as_uint(P.normal_mat[idx].w) : __float_as_uint(P.spheres[2 * (idx - P.ntri) + 1].x);
                                const float4 ks_tr = P.materials[4 * m + 2];
                                const uint32_t flags = __float_as_uint(P.materials[4 * m + 3].x);
                                lit = (flags & RT_HAS_TR) && (ks_tr.w < 1.0f);
                            }
                        }
                        if (!lit) P.lit[s] &= ~(1u << l);
                    }
                    __syncthreads();
                }
            }
            if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&P.counters[kCntExact]), (unsigned long long)nhit * P.nlights * (unsigned long long)P.ntri);
        }
        __syncthreads();
        // (4) shading + continuation rays
        shade_hits(P, level, tid, T);
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------
// Thread pencils (rt_tpencil.h): records, grouping, scan.
// ------------------------------------------------------------------------------------------------
// Records, position by position next to the generic ones ("never" there -- padding, degenerate, always-exact -- is "never" here).
__global__ void k_build_trec(const float4* __restrict__ triv, const float4* __restrict__ rec, int npos, int c1_end, int c2_end, float M, const TpSetup S,
                             float4* __restrict__ trec) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const float4 g3 = rec[4 * pos + 3];
    float q[24];
    tp_never(q);
    if (g3.x != kBminNever) {
        const uint32_t i = __float_as_uint(g3.y);
        const int W = pos < c1_end ? 0 : (pos < c2_end ? 1 : 2);
        const float4 A = triv[3 * i], B = triv[3 * i + 1], C = triv[3 * i + 2];
        const float a3f[3] = {A.x, A.y, A.z}, b3f[3] = {B.x, B.y, B.z}, c3f[3] = {C.x, C.y, C.z};
        const FilterTol t = filter_tolerances(a3f, b3f, c3f, W, (double)M);
        if (!t.always) tp_record(a3f, b3f, c3f, t.E0, t.E1, S, q);
    }
    q[7] = g3.y;    // id
    q[11] = g3.z;   // records in use in the tile (first record of a tile)
#pragma unroll
    for (int v = 0; v < kTpVec; ++v) trec[kTpVec * pos + v] = make_float4(q[4 * v], q[4 * v + 1], q[4 * v + 2], q[4 * v + 3]);
}

// One CTA: hist[t] accepted rays of reflector t -> groups of kTpR; exclusive scan of the group counts -> off[t];
// counters[kCntTpGroups] = number of groups.  (hist keeps the ray counts for k_tp_scatter, which also names each group's reflector.)
__global__ void __launch_bounds__(1024) k_tp_offsets(const __grid_constant__ FrameParams P) {
    __shared__ uint32_t part[1024];
    const int tid = threadIdx.x, T = blockDim.x;
    const int per = (P.ntri + T - 1) / T;
    const int t0 = tid * per, t1 = min(P.ntri, t0 + per);
    uint32_t sum = 0;
    for (int t = t0; t < t1; ++t) sum += P.tp_hist[t] / (uint32_t)kTpR;
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int i = 0; i < T; ++i) { const uint32_t v = part[i]; part[i] = run; run += v; }
        P.counters[kCntTpGroups] = run;
    }
    __syncthreads();
    uint32_t run = part[tid];
    for (int t = t0; t < t1; ++t) {
        P.tp_off[t] = run;
        run += P.tp_hist[t] / (uint32_t)kTpR;
    }
}

// Pool -> grouped queue.  A reflector's rays beyond its last full group go to the ordinary level-1 queue (generic scan).
__global__ void __launch_bounds__(256) k_tp_scatter(const __grid_constant__ FrameParams P) {
    const uint32_t count = P.counters[kCntTpPool];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (count + stride - 1) / stride;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t i = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool spill = false;
        uint32_t s = 0;
        if (i < count) {
            s = P.tp_pool[i];
            const uint32_t t = __float_as_uint(P.hit[s].w);   // the level-0 hit record is still in place: the reflector
            const uint32_t pos = atomicAdd(&P.tp_cursor[t], 1u);
            const uint32_t lim = (P.tp_hist[t] / (uint32_t)kTpR) * (uint32_t)kTpR;
            if (pos < lim) {
                P.q_tp[(size_t)P.tp_off[t] * kTpR + pos] = s;
                if ((pos % (uint32_t)kTpR) == 0) P.tp_group_tri[P.tp_off[t] + pos / (uint32_t)kTpR] = t;   // the group's first ray names its reflector
            } else spill = true;
        }
        warp_append(spill, s, P.q_ray, &P.counters[kCntRay + 1]);
    }
}

// k_trace_tp: nearest hit of the grouped level-1 rays.  One thread = one group = kTpR rays sharing the mirror image of the eye about
// their reflector's plane.  Same persistent CTAs, TMA ring (96-byte records), work items, cold state in shared memory and key
// merge as k_trace; scalar hot loop: per triangle 21 FMAs + 9 sign flips build the thread's oriented weight vectors, per ray
// 9 FMAs + LOP3s; the cold path rebuilds the block's mask with the full test (distance clauses, grazing clause of near planes).
struct TpSmem {
    KernelSmem<kTpR, false, kTpVec> ks;
    float lam_o[kTpR][kThreads], lam_hi[kTpR][kThreads];   // per-ray state only the cold path reads
};
template <int J, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_trace_tp(const __grid_constant__ FrameParams P, int level) {
    constexpr int R = kTpR;
    static_assert(R * J <= 32, "candidate mask must fit 32 bits");
    TpSmem& tsm = *reinterpret_cast<TpSmem*>(rt_dyn_smem);
    KernelSmem<R, false, kTpVec>& ks = tsm.ks;
    SmemCol<float> lam_o{tsm.lam_o}, lam_hi{tsm.lam_hi};
    const uint32_t ngroups = P.counters[kCntTpGroups];
    const Split sp = make_split(ngroups, kThreads, P.ntiles, true);
    Pipe pipe;
    pipe_init<false, kTpVec>(pipe, ks.sm, P.trec, sp, nullptr);
    uint32_t n_exact = 0;
    for (uint32_t item = blockIdx.x; item < sp.nitems; item += gridDim.x) {
        const uint32_t chunk = item / sp.parts, part = item - chunk * sp.parts;
        ColdState<R, true> st(ks.cold.get());
        auto& dist = st.dist; auto& best = st.best; auto& sid = st.sid;
        const uint32_t g = chunk * kThreads + threadIdx.x;
        const bool have = g < ngroups;
        float wx[R], wy[R], wz[R];
        float E[3] = {0.f, 0.f, 0.f};
        uint32_t live = 0;
        bool force_all = false;          // (cannot happen: the same function accepted these rays in k_shade) -> every triangle exact
        float lam_min_thread = FLT_MAX;
        if (have) {
            const uint32_t t = P.tp_group_tri[g];
            const float4 A = P.triv[3 * t], B = P.triv[3 * t + 1], C = P.triv[3 * t + 2];
            const float Af[3] = {A.x, A.y, A.z}, Bf[3] = {B.x, B.y, B.z}, Cf[3] = {C.x, C.y, C.z};
            if (!tp_mirror_point(P.tp.eye, P.tp.centerf, Af, Bf, Cf, E)) force_all = true;
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            sid[k] = 0; dist[k] = FLT_MAX; best[k] = -1; lam_hi[k] = FLT_MAX; lam_o[k] = 0.f;
            wx[k] = wy[k] = wz[k] = 0.f;
            if (have) {
                const uint32_t s = P.q_tp[(size_t)g * R + k];
                sid[k] = s;
                const float4 o = P.ray_o[s], d = P.ray_d[s];
                const float Of[3] = {o.x, o.y, o.z}, Df[3] = {d.x, d.y, d.z};
                TpRay tr;
                if (!tp_accepts(P.tp, E, Of, Df, tr)) force_all = true;
                wx[k] = tr.wx; wy[k] = tr.wy; wz[k] = tr.wz; lam_o[k] = tr.lam_o;
                lam_min_thread = fminf(lam_min_thread, tr.lam_o);
                live |= 1u << k;
            }
        }
        float cg = 1.0f, near_thr = 0.0f;
        if (have) tp_thread_consts(P.tp, lam_min_thread, cg, near_thr);
        if (force_all) near_thr = __int_as_float(0x7f800000);   // every plane is "near", and ...
        if (force_all) cg = __int_as_float(0x7f800000);         // ... every pair "grazing": all candidates
        const uint32_t deadmask = have ? 0u : 0x80000000u;
        const int tile_begin = (int)(part * sp.len);
        for (int tile = tile_begin; tile < tile_begin + (int)pipe.len; ++tile) {
            const float4* rec = pipe_acquire(pipe);
            if (__any_sync(0xffffffffu, live != 0u)) {
                const int nvalid = __float_as_int(rec[2].w);
#pragma unroll 1
                for (int jb = 0; jb < nvalid; jb += J) {
                    uint32_t acc[R];
#pragma unroll
                    for (int k = 0; k < R; ++k) acc[k] = 0xffffffffu;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        float q[24];
#pragma unroll
                        for (int v = 0; v < kTpVec; ++v) {
                            const float4 f = rec[(jb + j) * kTpVec + v];
                            q[4 * v] = f.x; q[4 * v + 1] = f.y; q[4 * v + 2] = f.z; q[4 * v + 3] = f.w;
                        }
                        TpTri T;
                        tp_orient(q, E, near_thr, T);
                        const uint32_t nearmask = T.near_ ? 0x7fffffffu : 0xffffffffu;
#pragma unroll
                        for (int k = 0; k < R; ++k) {
                            TpRay tr; tr.wx = wx[k]; tr.wy = wy[k]; tr.wz = wz[k]; tr.lam_o = 0.f;
                            acc[k] &= tp_weights(T, tr) & nearmask;
                        }
                    }
                    uint32_t all = 0xffffffffu;
#pragma unroll
                    for (int k = 0; k < R; ++k) all &= acc[k];
                    all |= deadmask;
                    if ((int)all >= 0) {   // cold: full test for the block, then exact re-evaluation per candidate pair
#pragma unroll 1
                        for (int j = 0; j < J; ++j) {
                            float q[24];
#pragma unroll
                            for (int v = 0; v < kTpVec; ++v) {
                                const float4 f = rec[(jb + j) * kTpVec + v];
                                q[4 * v] = f.x; q[4 * v + 1] = f.y; q[4 * v + 2] = f.z; q[4 * v + 3] = f.w;
                            }
                            TpTri T;
                            tp_orient(q, E, near_thr, T);
                            const int tri = (int)__float_as_uint(q[7]);
#pragma unroll
                            for (int k = 0; k < R; ++k) {   // (unrolled: the ray registers must not be indexed dynamically)
                                if (!((live >> k) & 1u)) continue;
                                TpRay tr; tr.wx = wx[k]; tr.wy = wy[k]; tr.wz = wz[k]; tr.lam_o = lam_o[k];
                                const float lam_lo = __fsub_rd(tr.lam_o, P.tp.lam_slack);
                                if (!tp_candidate(T, tr, lam_hi[k], lam_lo, cg)) continue;
                                const float4 o = P.ray_o[sid[k]], d = P.ray_d[sid[k]];
                                const float4 e = exact_eval_tri(P.triv, tri, o.x, o.y, o.z, d.x, d.y, d.z);
                                ++n_exact;
                                if (!(e.w < 0.0f) && (e.w < dist[k] || (e.w == dist[k] && tri < best[k]))) {   // (distance, id) order: raytracing.cpp:183
                                    dist[k] = e.w;
                                    best[k] = tri;
                                    const float h = __fadd_ru(tr.lam_o, __fadd_ru(e.w, P.tp.lam_slack));
                                    lam_hi[k] = (h < FLT_MAX) ? h : FLT_MAX;
                                }
                            }
                        }
                    }
                }
            }
            pipe_release<false>(pipe);
        }
#pragma unroll
        for (int k = 0; k < R; ++k)
            if (((live >> k) & 1u) && best[k] >= 0)
                atomicMin(&P.key[sid[k]], ((unsigned long long)__float_as_uint(dist[k]) << 32) | (unsigned int)best[k]);
    }
    if (n_exact) atomicAdd(reinterpret_cast<unsigned long long*>(&P.counters[kCntExact]), (unsigned long long)n_exact);
}

// ------------------------------------------------------------------------------------------------
// k_resolve: per pixel, sum the samples (subx outer, suby inner == ascending sample index), divide by
// raysPerPixel with a true division (Vec3D.h:36-38), clamp like RGBValue (main.cpp:29-41).
// ------------------------------------------------------------------------------------------------
__global__ void k_resolve(const __grid_constant__ FrameParams P, float* __restrict__ fb_local) {
    const uint32_t npix = P.nrows * P.W;
    const uint32_t spp = P.pfx * P.pfy;
    const float rays = (float)(int)spp;
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
        float r = 0.f, g = 0.f, b = 0.f;
        for (uint32_t k = 0; k < spp; ++k) {
            const float4 a = P.acc[pix * spp + k];
            r = __fadd_rn(r, a.x); g = __fadd_rn(g, a.y); b = __fadd_rn(b, a.z);
        }
        float ch[3] = {__fdiv_rn(r, rays), __fdiv_rn(g, rays), __fdiv_rn(b, rays)};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (ch[c] > 1.0f) ch[c] = 1.0f;
            if (ch[c] < 0.0f) ch[c] = 0.0f;   // NaN passes both, like the reference
        }
        float* dst = fb_local + 3 * ((size_t)P.row0 * P.W + pix);
        dst[0] = ch[0]; dst[1] = ch[1]; dst[2] = ch[2];
    }
}

// gathered: [G][rows_per_rank][W][3] (rank-major, as ncclAllGather leaves it) -> final [H][W][3]
__global__ void k_deinterleave(const float* __restrict__ gathered, float* __restrict__ final_fb, uint32_t W, uint32_t H, uint32_t G, uint32_t rows_per_rank) {
    const size_t n = (size_t)W * H * 3;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / (3u * W));
        const uint32_t rem = (uint32_t)(i - (size_t)y * 3u * W);
        const uint32_t rank = y % G, ly = y / G;
        final_fb[i] = gathered[((size_t)rank * rows_per_rank + ly) * 3u * W + rem];
    }
}

// Detached rank (world > 1 without a communicator, used to test the row interleave on one device): this
// rank's rows go to their final position, every other row is zero.
__global__ void k_place_rows(const float* __restrict__ local, float* __restrict__ final_fb, uint32_t W, uint32_t H, uint32_t G, uint32_t rank) {
    const size_t n = (size_t)W * H * 3;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / (3u * W));
        const uint32_t rem = (uint32_t)(i - (size_t)y * 3u * W);
        final_fb[i] = (y % G == rank) ? local[(size_t)(y / G) * 3u * W + rem] : 0.0f;
    }
}

// rt_trace: the caller's rays arrive packed, (origin, dest) = 2 float4 per ray; sets up the per-sample wavefront state
// the primary scan of a frame would have written (level 0, throughput 1, colour 0).
__global__ void k_init_trace(const float4* __restrict__ rays, int n, float4* __restrict__ ray_o, float4* __restrict__ ray_d,
                             float4* __restrict__ thr, float4* __restrict__ acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = rays[2 * i], t = rays[2 * i + 1];
    ray_o[i] = make_float4(o.x, o.y, o.z, 0.f);
    ray_d[i] = make_float4(t.x, t.y, t.z, __int_as_float(0));
    thr[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Measured FP32 ceiling of the device (rt_probe_fp32_peak): 16 independent packed-FMA chains per thread whose
// multiplier and addend stay in the operand-reuse cache, i.e. the most the FMA pipe can retire (2 x 2 flop per
// FFMA2 per lane).  Not part of the render path.
template <bool PACKED>
__global__ void __launch_bounds__(256) k_fp32_peak_probe(float2* __restrict__ out, float seed, int iters) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(seed + i, seed - i + threadIdx.x * 1e-3f);
    const float2 m = make_float2(0.9999f, 1.0001f), c = make_float2(1e-4f, -1e-4f);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (PACKED) acc[i] = __ffma2_rn(acc[i], m, c);                                   // 1 FFMA2
            else { acc[i].x = fmaf(acc[i].x, m.x, c.x); acc[i].y = fmaf(acc[i].y, m.y, c.y); }  // 2 FFMA
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Image::writeImage's quantiser (main.cpp:117): (unsigned char)(v * 255.0f), truncation toward zero.
__global__ void k_quantise(const float* __restrict__ fb, uint8_t* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = __fmul_rn(fb[i], 255.0f);
        out[i] = (uint8_t)(int)v;  // values are clamped to [0,1] upstream, so int conversion == uchar conversion
    }
}

}  // namespace rt
