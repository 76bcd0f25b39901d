/* rt_b200.h -- C ABI of librt_b200.so, the B200-native replacement for the render hot path of the
 * TU Delft TI1805 ray tracer (wmorssink/raytracert).
 *
 * The reference has no FFI: its hot path sits behind the course's C++ "plugin" contract
 * (CG_Project/raytracing.h:8-41) and is driven by the 'r' key handler (CG_Project/main.cpp:340-412).
 * Each entry point below names the reference interface it replaces.  Plain pointers and sizes only:
 * no C++ types, no torch types.  There is NO CPU fallback and no backend dispatch: if no CUDA device
 * is usable every call returns RT_ERR_NO_DEVICE (see rt_last_error()).
 *
 * Threading: call from one host thread. The library drives all of its devices internally.
 * Errors: 0 = ok, negative = error code, text via rt_last_error(); nothing is thrown across the ABI.
 * Ownership: the library copies what it is given; host buffers stay owned by the caller.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_OK 0
#define RT_ERR_NO_DEVICE (-1)    /* no usable CUDA device / driver (there is no CPU path)          */
#define RT_ERR_INVALID (-2)      /* bad argument                                                   */
#define RT_ERR_STATE (-3)        /* call order (e.g. rt_render before rt_upload_scene)             */
#define RT_ERR_CUDA (-4)         /* a CUDA runtime call failed                                     */
#define RT_ERR_NCCL (-5)         /* NCCL missing or a collective failed                            */

/* Feature toggles == the reference's global bools (raytracing.cpp:15-20, keys '1'..'6' :456-473). */
#define RT_AMBIENT 1u
#define RT_DIFFUSE 2u
#define RT_SPECULAR 4u
#define RT_REFLECTION 8u
#define RT_SHADOWS 16u
#define RT_REFRACTION 32u
#define RT_ALL_FEATURES 63u

/* Material flags == Material::has_*() (mesh.h:58-64). */
#define RT_HAS_KD 1u
#define RT_HAS_KA 2u
#define RT_HAS_KS 4u
#define RT_HAS_NS 8u
#define RT_HAS_NI 16u
#define RT_HAS_TR 32u

#define RT_MAX_LIGHTS 16 /* the reference's 'L' key (main.cpp:334-336) has no limit: more lights are an error here (RT_ERR_INVALID), never a truncated frame */

/* One Material (mesh.h:116-122), 64 bytes, laid out as four float4 for the device. */
typedef struct rt_material {
    float Kd[3], Ns;
    float Ka[3], Ni;
    float Ks[3], Tr;
    uint32_t flags; /* RT_HAS_* */
    uint32_t pad[3];
} rt_material;

/* Analytic sphere primitive (Sphere.h:14-32: center, radius, material). */
typedef struct rt_sphere {
    float center[3], radius;
    uint32_t material;
    uint32_t pad[3];
} rt_sphere;

/* The scene as Mesh (mesh.h:172-201) + the face-normal table (raytracing.cpp:33,78-86) flattened to SoA
 * float4 arrays: triangle i has corners v0[i], v1[i], v2[i] (xyz, w ignored) in the order of
 * Mesh::triangles[i].v[0..2], unit face normal normal[i] (xyz, w ignored) and material tri_material[i]
 * (index into materials). All float4 arrays are n_triangles * 4 floats. */
typedef struct rt_scene {
    uint32_t n_triangles;
    const float* v0;
    const float* v1;
    const float* v2;
    const float* normal;
    const uint32_t* tri_material;
    uint32_t n_materials;
    const rt_material* materials;
    uint32_t n_spheres; /* may be 0 */
    const rt_sphere* spheres;
} rt_scene;

/* Everything the 'r' handler reads when it renders a frame (main.cpp:347-362 and the globals of
 * raytracing.h:8-16 / raytracing.cpp:15-29). */
typedef struct rt_params {
    /* produceRay() (main.cpp:300-320) at the four corners, in the order of main.cpp:355-358:
     * (0,0) (0,H-1) (W-1,0) (W-1,H-1), each as origin xyz then dest xyz -> 24 floats. */
    float corners[24];
    uint32_t width, height;               /* WindowSize_X / WindowSize_Y                        */
    uint32_t pixelfactor_x, pixelfactor_y; /* rays per pixel = pfx*pfy (raytracing.cpp:23-25)    */
    int32_t max_lvl;                       /* recursion bound (raytracing.cpp:29)                */
    uint32_t features;                     /* RT_AMBIENT | ...                                   */
    float camera[3];                       /* MyCameraPosition (used by the specular term)       */
    uint32_t n_lights;                     /* MyLightPositions                                    */
    float lights[RT_MAX_LIGHTS][3];
    uint32_t want_prim_id;                 /* also keep the per-sample primary primitive id      */
} rt_params;

/* Counters of the last completed rt_render / rt_trace (this process's share of the frame). */
typedef struct rt_stats {
    uint64_t primary_rays, shadow_rays, bounce_rays; /* == intersectMesh calls by kind              */
    uint64_t tri_tests;                               /* rays * triangles, the reference's count     */
    uint64_t exact_evals;                             /* (ray,triangle) pairs re-done in exact order */
    /* device time (CUDA events on the library's stream; max over this process's GPUs): the whole frame, and the
     * sum over launches of each kernel kind: nearest-hit scans (k_trace), shadow scans (k_shadow), hit records +
     * shading (k_finish, k_shade), resolve, and the all-gather + de-interleave */
    float ms_total, ms_trace, ms_shadow, ms_shade, ms_resolve, ms_gather;
    uint32_t n_gpus, rank, n_triangles, n_levels;
    uint32_t n_launches;                              /* kernels launched for the frame (per GPU)      */
    uint32_t variant;                                 /* bit 0: scan kernels without the grazing clause (fine mesh);
                                                       * bit 1: pencil filter on the primary rays; bit 2: on shadow rays;
                                                       * bit 3: pencil records built without the clause-free proof;
                                                       * bit 4: the frame / batch was a CUDA-graph replay;
                                                       * bit 5: reflection (mirror) pencils served level-1 continuation rays;
                                                       * bit 6: thread pencils did */
    float ms_trace_primary;                           /* the level-0 (primary ray) part of ms_trace */
    float ms_trace_mirror;                            /* the part of ms_trace spent in mirror-pencil scans (RT_OPT_PENCIL_REFLECT) */
    uint64_t mirror_rays;                             /* level-1 continuation rays served by a mirror pencil (part of bounce_rays) */
    uint64_t thread_pencil_rays;                      /* level-1 continuation rays served by thread pencils (part of bounce_rays) */
    float ms_trace_thread;                            /* the part of ms_trace spent in thread-pencil scans (RT_OPT_PENCIL_THREAD) */
    uint32_t reserved;
} rt_stats;

/* Single-process mode: use devices 0..n_gpus-1 of this box (n_gpus >= 1); rows are interleaved over
 * them and gathered with one NCCL all-gather per frame (ncclCommInitAll). */
int rt_init(int n_gpus);

/* One-process-per-GPU mode (torchrun): this process drives `device` as rank `rank` of `world`.
 * nccl_id: the bytes of an ncclUniqueId made by rank 0 with rt_nccl_unique_id() and distributed by the
 * caller (ignored when world == 1).  nccl_id == NULL with world > 1 makes a DETACHED rank: it renders its
 * own rows (y % world == rank) and rt_download_framebuffer returns them in place with every other row
 * zero -- no communicator, no all-gather (used to test the row interleave on a single device). */
int rt_init_rank(int device, int rank, int world, const void* nccl_id, size_t nccl_id_bytes);
int rt_nccl_unique_id(void* out, size_t cap, size_t* bytes);

/* replaces: init() tail (raytracing.cpp:63-67) -- Mesh + normals become device-resident buffers. */
int rt_upload_scene(const rt_scene* scene);

/* replaces: the y/x/subx/suby loop of main.cpp:369-395 incl. performRayTracing() per sample and the
 * RGBValue clamp. Returns when the frame (and the all-gather) has completed. */
int rt_render(const rt_params* params);
/* Same work, enqueue only (for back-to-back timing); pair with rt_sync(). */
int rt_render_async(const rt_params* params);
int rt_sync(void);

/* replaces: Image::setPixel target (main.cpp:88-94). rgb: 3*W*H floats, row 0 first, clamped to [0,1].
 * prim_id (optional, needs want_prim_id): W*H*pfx*pfy ints, sample ((y*W+x)*pfx+subx)*pfy+suby, -1 = miss. */
int rt_download_framebuffer(float* rgb, int32_t* prim_id);
/* replaces: the quantiser of Image::writeImage (main.cpp:116-117): (unsigned char)(v*255.0f). */
int rt_download_framebuffer_u8(uint8_t* rgb8);

/* replaces: performRayTracing(origin, dest) (raytracing.cpp:410-416) for a batch of n rays; uses the
 * camera/lights/toggles/max_lvl of `params` (corners/size ignored). origins/dests: 3*n floats.
 * rgb: 3*n floats (unclamped). prim_id / hit (optional): nearest primitive and its intersection point. */
int rt_trace(const rt_params* params, int n, const float* origins, const float* dests, float* rgb,
             int32_t* prim_id, float* hit);

/* Options (rt_set_option; they persist until rt_shutdown -- across rt_init, so they may be set before it).
 * RT_OPT_TILE_CULLING (default 0): 1 = conservative tile culling.  Triangles are scanned in tiles of 128; with this
 *   option only tiles (found through a two-level hierarchy of bounding boxes) that some ray of a thread block can reach
 *   are streamed, and a warp skips a tile when none of its own rays can reach the box of the tile's (tolerance-dilated)
 *   triangles.  Scenes of more than 2.1 M triangles are scanned brute force.  The image and the primitive ids are IDENTICAL to the brute-force scan (same filter + exact tiers on every
 *   tile that is not skipped) -- it only stops being the O(rays x triangles) loop of the reference
 *   (raytracing.cpp:174-189), which is why it is opt-in and reported separately by bench.py.  Set it BEFORE
 *   rt_upload_scene to also get spatially sorted tiles (Morton order of the triangle centroids), which makes the boxes
 *   compact; enabling it afterwards works on the tiles in file order. */
#define RT_OPT_TILE_CULLING 1
/* RT_OPT_PENCIL (default 1): 1 = primary rays (all lines pass through the eye) and the shadow rays of a light (all end at
 * the light) are filtered with the common-point ("pencil") form of the ray-triangle test -- 12 instead of 16 packed FP32
 * instructions per (ray pair, triangle) -- whenever the frame qualifies (brute-force scan, clause-free scene, perspective
 * camera / light outside the scene box; rt_stats.variant says what was used).  Same filter + exact tiers, identical
 * image and ids.  0 = always the generic filter. */
#define RT_OPT_PENCIL 2
/* RT_OPT_PENCIL_ANY (default 1): the pencil filter is also used for scenes without the scene-level clause-free proof
 * (large triangles: cube.obj, dodgeColorTest.obj).  Triangles whose plane passes within lam_max*cos_g + 2*delta of the
 * common point get "always candidate" records -- the exact path decides for every ray -- (at most 16 per launch, else
 * that launch keeps the generic kernels); rt_stats.variant bit 3 says it was used.  0 = pencil launches only under the
 * clause-free proof.  Same filter + exact tiers, identical image and ids (tests/test_gpu_parity.py). */
#define RT_OPT_PENCIL_ANY 3
/* RT_OPT_PENCIL_REFLECT (default 1): the continuation rays of PRIMARY hits on a large group of coplanar triangles (a floor,
 * a wall, water) are the mirror image of a pencil through the eye (reflection(), raytracing.cpp:277-285): they are filtered
 * with pencil records around the mirrored eye instead of the generic filter.  Every such ray is checked against its pencil
 * (distance of its line to the mirrored eye, chart, origin side) when it is spawned; a ray that fails takes the generic scan.
 * At most 4 plane groups per scene; rt_stats.variant bit 5, rt_stats.mirror_rays.  Identical image and ids. */
#define RT_OPT_PENCIL_REFLECT 5
/* RT_OPT_GRAPH (default -1 = auto): small frames and small rt_trace batches (samples x triangles <= 4e9: the launch gaps
 * would dominate -- cube.obj at 800x800 is 44 launches for 0.6 ms) are replayed from a captured CUDA graph (every scan
 * launch has a fixed grid: persistent CTAs read their ray counts from device counters); rt_stats.variant bit 4.
 * 0 = never, 1 = always.  A replayed frame reports no per-kernel times (rt_stats.ms_trace .. ms_resolve are 0). */
#define RT_OPT_GRAPH 4
/* RT_OPT_PENCIL_THREAD (default 1 = auto: frames of more than 4e9 sample-triangle pairs; 2 = always; 0 = never): the level-1 continuation rays of the primary hits on ANY triangle leave that triangle's own
 * mirror image of the eye; grouped by reflector (8 rays per thread), they are scanned with per-thread pencil weights built on
 * the fly from an E-independent record (rt_tpencil.h) instead of the generic filter.  Every ray is checked against its
 * pencil when it is spawned; rays that do not fill a group take the generic scan.  rt_stats.variant bit 6,
 * rt_stats.thread_pencil_rays.  Identical image and ids. */
#define RT_OPT_PENCIL_THREAD 7
/* RT_OPT_SMALL_TRACE (default 1): rt_trace batches of at most 32 rays (and at most 1e5 ray-triangle pairs per level) -- the
 * drop-in performRayTracing(origin, dest) call is a batch of one -- run the whole recursion in ONE kernel launch of one
 * thread block with exact tests only (k_trace_small) instead of four launches per level.  0 = always the wavefront kernels. */
#define RT_OPT_SMALL_TRACE 6
int rt_set_option(int option, int value);

int rt_get_stats(rt_stats* out);

/* Measurement aid, not on the render path: the FP32 rate (TFLOP/s, FMA = 2) device 0 sustains on a register-resident
 * packed-FMA loop -- the practical ceiling the roofline fraction of the intersection kernels can be read against. */
int rt_probe_fp32_peak(float* tflops);

/* Device-side timing on the library's own stream (CUDA events): slots 0..15. */
int rt_event_record(int slot);
int rt_event_elapsed_ms(int slot_begin, int slot_end, float* ms);

const char* rt_last_error(void);
void rt_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
