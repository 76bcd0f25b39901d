// rt_b200.cu -- librt_b200.so: contexts, scene upload, frame orchestration and the C ABI of
// include/rt_b200.h.  One RtDevice per GPU this process drives; rt_init(n) drives n GPUs from one process
// (the C++ drop-in, ncclCommInitAll), rt_init_rank() drives one GPU as a rank of a torchrun job.
// No CPU fallback: without a CUDA device every entry point fails with RT_ERR_NO_DEVICE.
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "rt_kernels.cuh"

namespace {

using namespace rt;

// ---- tunables -------------------------------------------------------------------------------------
// Scan-kernel shapes compiled into the library: (ray pairs per thread, triangles per filter block,
// resident CTAs per SM).  The first entry is the default; RT_B200_TUNE="rp,j,minb" selects another
// (tools/tune.py sweeps them on the GPU).
#define RT_SCAN_CONFIGS(X) X(2, 8, 2) X(1, 16, 4)
// The pencil kernels keep 6 registers per ray pair instead of 14, so they have their own shapes (RT_B200_PTUNE="rp,j,minb").
#define RT_PENCIL_CONFIGS(X) X(2, 8, 2) X(1, 16, 4) X(4, 4, 2) X(3, 4, 2)   // rp >= 3: scalar hot loop (rt_kernels.cuh: pencil_ray_hot), cold state in dynamic shared memory
struct ScanConfig { int rp, j, minb; };
constexpr double kGraphMaxTests = 4e9;     // samples x triangles below which RT_OPT_GRAPH = auto captures the frame (and thread pencils stay off)
constexpr uint32_t kMaxChunkSamples = 1u << 23;  // 8 Mi samples per wavefront chunk (84 B of state each)

// ---- NCCL through dlopen (no link-time dependency; inside python the already-loaded torch copy is reused)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) return false;
#define RT_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(handle, name)); if (!field) return false;
        RT_SYM(GetUniqueId, "ncclGetUniqueId")
        RT_SYM(CommInitRank, "ncclCommInitRank")
        RT_SYM(CommInitAll, "ncclCommInitAll")
        RT_SYM(CommDestroy, "ncclCommDestroy")
        RT_SYM(AllGather, "ncclAllGather")
        RT_SYM(GroupStart, "ncclGroupStart")
        RT_SYM(GroupEnd, "ncclGroupEnd")
        RT_SYM(GetErrorString, "ncclGetErrorString")
#undef RT_SYM
        return true;
    }
};
constexpr int kNcclFloat = 7;  // ncclFloat32

// ---- state ------------------------------------------------------------------------------------------
struct RtDevice {
    int device = 0, rank = 0;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    // scene
    float4 *rec = nullptr, *triv = nullptr, *normal_mat = nullptr, *materials = nullptr, *spheres = nullptr, *tile_box = nullptr, *super_box = nullptr;
    size_t cap_box = 0, cap_super = 0;
    size_t cap_rec = 0, cap_triv = 0, cap_nm = 0, cap_mat = 0, cap_sph = 0;  // in float4; buffers are reused across uploads
    int ntri = 0, ntiles = 0, nmat = 0, nspheres = 0;
    int cls1 = 0, cls2 = 0;             // first tile of dominant-axis class 1 / 2
    uint32_t* perm = nullptr;           // record position -> triangle id (kNoTriangle = padding)
    size_t cap_perm = 0;
    float M_built = 0.f, dir_built = 0.f;   // magnitude bound / longest ray the records were built for (0: not built)
    uint64_t rec_gen = 0;                   // bumped by every (re)build of the records: invalidates cached pencil records / graphs
    float* h_small = nullptr;               // pinned: n_always + scene box read back by build_records, n_near by plan_pencil
    cudaEvent_t ev_stage = nullptr;         // recorded after the last H2D copy out of the upload staging buffers
    std::vector<uint8_t> plan_key;          // what the cached pencil records were built from (plan_pencil)
    std::vector<uint8_t> plan_blob;         // the cached PencilPlan
    cudaGraphExec_t frame_graph = nullptr;  // captured wavefront of a small frame (RT_OPT_GRAPH), valid while graph_key matches
    std::vector<uint8_t> graph_key;
    bool capturing = false;                 // launches go into a stream capture: no per-launch events
    uint32_t launches_in_graph = 0;
    int levels_in_graph = 0;
    bool used_graph = false;                // the last frame was a graph replay (rt_stats.variant bit 4)
    bool no_grazing = false;                // the records carry no grazing clause (see build_records)
    unsigned int* n_always = nullptr;       // device counter written by k_build_records
    uint32_t* always_list = nullptr;        // triangles outside the filter (see k_build_records)
    size_t cap_always = 0;
    int n_always_host = 0;
    uint32_t pencil_used = 0;               // rt_stats.variant bits of the last frame (2: primary rays, 4: shadow rays)
    // pencil filter (rt_pencil.h): slot 0 = records around the eye, slot 1 + l = around light l; rebuilt every frame
    float4* prec = nullptr; size_t cap_prec = 0;
    // reflection pencils: plane group of every triangle, ray queues of the groups (kMaxMirrors x cap_samples)
    // thread pencils (rt_tpencil.h): records (rebuilt with the pencil plan), pool / grouped queue (per chunk), per-triangle tables
    float4* trec = nullptr; size_t cap_trec = 0;
    uint32_t *tp_pool = nullptr, *q_tp = nullptr, *tp_group_tri = nullptr; size_t cap_tp = 0;
    uint32_t *tp_hist = nullptr, *tp_off = nullptr, *tp_cursor = nullptr; size_t cap_tp_tri = 0;
    uint8_t* tri_group = nullptr; size_t cap_group = 0;
    uint32_t* q_mirror = nullptr; size_t cap_q_mirror = 0;
    float4* scene_box = nullptr;            // device: union of the tile boxes (k_scene_box)
    unsigned int* n_near = nullptr;         // device: "always candidate" records per pencil slot (RT_OPT_PENCIL_ANY)
    float box_lo[3] = {0, 0, 0}, box_hi[3] = {0, 0, 0};   // host copy, valid while the generic records are
    // per-chunk state
    size_t cap_samples = 0;
    float4 *ray_o = nullptr, *ray_d = nullptr, *thr = nullptr, *acc = nullptr, *hit = nullptr;
    uint32_t *lit = nullptr, *q_ray = nullptr, *q_hit = nullptr;
    unsigned long long* key = nullptr;  // nearest-hit merge keys, kKeyEmpty between launches
    bool key_dirty = false;             // a call failed between a scan and its k_finish: re-initialise before the next use
    float4* hit0 = nullptr; size_t cap_hit0 = 0;          // rt_trace: copy of the level-0 hit records (grow-only)
    float4* trace_in = nullptr; size_t cap_trace_in = 0;  // rt_trace: packed (origin, dest) rays as uploaded
    cudaGraphExec_t trace_graph = nullptr;                // rt_trace: captured copies + wavefront of a small batch
    std::vector<uint8_t> trace_key;
    uint32_t launches_in_trace_graph = 0;
    uint32_t* counters = nullptr;      // kCntWords per chunk slot
    int counters_slots = 0;
    int frame_chunks = 0;              // counter slots the last frame / trace call used
    int32_t* prim = nullptr; size_t cap_prim = 0;
    // framebuffers
    float *fb_local = nullptr, *fb_gather = nullptr, *fb_final = nullptr;
    size_t cap_local = 0, cap_gather = 0, cap_final = 0;
    uint8_t* fb_u8 = nullptr; size_t cap_u8 = 0;
    cudaEvent_t ev[16] = {};
    cudaEvent_t ev_phase[8] = {};
    // per-launch timing: events 2*i / 2*i+1 bracket launch i of the current frame; kind: see KernelKind
    std::vector<cudaEvent_t> kev;
    std::vector<int> kev_kind;
    int num_sms = 148;
};

enum KernelKind { kKindTrace = 0, kKindShadow, kKindShade, kKindResolve, kKindGather, kKindTracePrimary, kKindTraceMirror, kKindTraceThread, kNumKinds };
constexpr size_t kMaxTimedLaunches = 4096;  // beyond this a frame's launches are still counted, not timed

struct Global {
    std::vector<RtDevice> devs;
    int world = 0;          // total ranks (== devs.size() in single-process mode)
    bool single_process = true;
    bool scene_ready = false, frame_ready = false;   // frame_ready: a framebuffer can be downloaded
    bool stats_ready = false;                        // the counters describe a completed rt_render / rt_trace
    NcclApi nccl;
    ScanConfig scan = {2, 8, 2};
    ScanConfig pscan = {2, 8, 2};    // shape of the pencil kernels
    bool tile_culling = false;       // RT_OPT_TILE_CULLING
    bool pencil = true;              // RT_OPT_PENCIL: common-point filter for primary / shadow rays where it applies
    int pencil_thread = 1;           // RT_OPT_PENCIL_THREAD (0 never, 1 auto: frames that are not launch-bound, 2 always): per-thread pencils for the level-1 continuation rays of every other triangle
    bool pencil_reflect = true;      // RT_OPT_PENCIL_REFLECT: mirror pencils for the level-1 continuation rays of planar reflectors
    struct PlaneGroup { double n[3], d; uint32_t count; };
    std::vector<PlaneGroup> planes;  // the (at most kMaxMirrors) largest groups of coplanar triangles of the scene; group id = index
    bool pencil_any = true;          // RT_OPT_PENCIL_ANY: also for scenes without the clause-free proof (near-plane triangles: always candidates)
    bool small_trace = true;         // RT_B200_SMALL_TRACE=0 keeps small rt_trace batches on the multi-launch wavefront (tests / A-B)
    int graph_mode = -1;             // RT_OPT_GRAPH: -1 auto (small frames replay a captured CUDA graph), 0 never, 1 always
    bool allow_no_grazing = true;    // RT_B200_GRAZING=1 forces the grazing clause on (experiments)
    float max_uv = 0.f;              // max over the triangles of |v1-v0| * |v2-v0| (see build_records)
    float max_ni = 1.f;              // max over the materials of max(Ni, 1/Ni): bounds the refracted direction
    bool unit_normals = false;       // no face normal handed to rt_upload_scene is longer than 1 (+ rounding)
    float cos_min = kCosMinDefault;  // grazing threshold of the filter (RT_B200_COSMIN overrides, for experiments)
    float scene_extent = 0.f;   // max |coordinate| over the scene
    bool any_transparent = false;
    rt_params last;             // params of the last frame
    rt_params last_trace;       // params of the last rt_trace call
    bool stats_of_trace = false;   // the counters describe an rt_trace call (of last_trace_n rays), not a frame
    int last_trace_n = 0;
    uint32_t rows_per_rank = 0;
    rt_stats stats = {};
    std::vector<uint32_t> host_counters;
    std::string error;
} g;

// Host staging for rt_upload_scene: page-locked, grow-only, reused across uploads, so the H2D copies are true DMA
// transfers from pinned memory.  Falls back to pageable memory if pinning fails.
template <class T>
struct PinnedStage {
    T* p = nullptr;
    size_t cap = 0, n = 0;
    bool pinned = false;
    void resize(size_t count) {
        if (count > cap) {
            release();
            const size_t want = count + count / 8 + 16;
            if (cudaHostAlloc((void**)&p, want * sizeof(T), cudaHostAllocDefault) == cudaSuccess) pinned = true;
            else { cudaGetLastError(); p = (T*)malloc(want * sizeof(T)); pinned = false; }
            cap = p ? want : 0;
        }
        n = count;
    }
    void assign(size_t count, const T& v) { resize(count); for (size_t i = 0; i < n; ++i) p[i] = v; }
    void release() { if (p) { if (pinned) cudaFreeHost(p); else free(p); } p = nullptr; cap = n = 0; }
    T* data() { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) { return p[i]; }
    T* begin() { return p; }
};
PinnedStage<float4> g_stage_triv, g_stage_nm, g_stage_sph;
PinnedStage<uint32_t> g_stage_perm;
PinnedStage<uint8_t> g_stage_group;
PinnedStage<rt_material> g_stage_mat;
PinnedStage<float4> g_stage_rays, g_stage_out;   // rt_trace: rays in (origin, dest), results out (colour, level-0 hit)

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g.error = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(RT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
#define NC(call)                                                                                                 \
    do {                                                                                                         \
        ncclResult_t r_ = (call);                                                                                \
        if (r_ != 0) return fail(RT_ERR_NCCL, "%s failed: %s", #call, g.nccl.GetErrorString ? g.nccl.GetErrorString(r_) : "?"); \
    } while (0)

template <class T>
int ensure(T*& ptr, size_t& cap, size_t need) {
    if (need <= cap && ptr) return RT_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    CU(cudaMalloc(&ptr, need * sizeof(T)));
    cap = need;
    return RT_OK;
}

int create_device(RtDevice& d, int device, int rank) {
    d.device = device; d.rank = rank;
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    for (auto& e : d.ev) CU(cudaEventCreate(&e));
    for (auto& e : d.ev_phase) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&d.ev_stage, cudaEventDisableTiming));
    CU(cudaHostAlloc((void**)&d.h_small, 64 * sizeof(float), cudaHostAllocDefault));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(RT_ERR_NO_DEVICE, "device %d is sm_%d%d; librt_b200 is built for sm_100a only", device, prop.major, prop.minor);
    d.num_sms = prop.multiProcessorCount;
    return RT_OK;
}

void destroy_device(RtDevice& d) {
    cudaSetDevice(d.device);
    if (d.comm && g.nccl.CommDestroy) g.nccl.CommDestroy(d.comm);
    void* ptrs[] = {d.trec, d.tp_pool, d.q_tp, d.tp_group_tri, d.tp_hist, d.tp_off, d.tp_cursor, d.tri_group, d.q_mirror, d.n_near, d.prec, d.scene_box, d.rec, d.perm, d.n_always, d.always_list, d.tile_box, d.super_box, d.triv, d.normal_mat, d.materials, d.spheres, d.ray_o, d.ray_d, d.thr, d.acc, d.hit, d.lit, d.q_ray,
                    d.q_hit, d.key, d.hit0, d.trace_in, d.counters, d.prim, d.fb_local, d.fb_gather, d.fb_final, d.fb_u8};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& e : d.ev) if (e) cudaEventDestroy(e);
    for (auto& e : d.ev_phase) if (e) cudaEventDestroy(e);
    for (auto& e : d.kev) cudaEventDestroy(e);
    if (d.frame_graph) cudaGraphExecDestroy(d.frame_graph);
    if (d.trace_graph) cudaGraphExecDestroy(d.trace_graph);
    if (d.ev_stage) cudaEventDestroy(d.ev_stage);
    if (d.h_small) cudaFreeHost(d.h_small);
    if (d.stream) cudaStreamDestroy(d.stream);
    d = RtDevice();
}

// Brackets one kernel launch with events on the device's stream (cheap: two event records).
struct LaunchTimer {
    RtDevice& d;
    bool timed;
    LaunchTimer(RtDevice& dev, int kind) : d(dev), timed(!dev.capturing && dev.kev_kind.size() < kMaxTimedLaunches) {
        if (timed) {
            const size_t i = d.kev_kind.size();
            while (d.kev.size() < 2 * (i + 1)) { cudaEvent_t e; cudaEventCreate(&e); d.kev.push_back(e); }
            cudaEventRecord(d.kev[2 * i], d.stream);
        }
        d.kev_kind.push_back(timed ? kind : -1 - kind);
    }
    ~LaunchTimer() { if (timed) cudaEventRecord(d.kev[2 * (d.kev_kind.size() - 1) + 1], d.stream); }
};

enum { kScanPrimary = 0, kScanBounce = 1, kScanShadowAny = 2, kScanShadowNearest = 3, kScanPrimaryPencil = 4, kScanShadowPencil = 5, kScanBouncePencil = 6 };

// Every scan kernel takes its shared memory (TMA ring, cold per-ray state, culling lists) as one dynamic block; the
// attribute is raised once per kernel instantiation and device (> 48 KB needs the opt-in).
template <class K>
void launch_kernel_smem(K kernel, size_t smem, int grid, cudaStream_t st, const FrameParams& P, int level) {
    // (K is the same function-pointer type for every kernel: the bookkeeping is keyed by the pointer's value)
    static std::vector<std::pair<const void*, uint64_t>> raised;
    int dev = 0;
    cudaGetDevice(&dev);
    const void* fn = reinterpret_cast<const void*>(kernel);
    uint64_t* mask = nullptr;
    for (auto& e : raised) if (e.first == fn) { mask = &e.second; break; }
    if (!mask) { raised.emplace_back(fn, 0ull); mask = &raised.back().second; }
    if (dev >= 0 && dev < 64 && !((*mask >> dev) & 1ull)) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        *mask |= 1ull << dev;
    }
    kernel<<<grid, kThreads, smem, st>>>(P, level);
}

template <int RP, int J, int MINB, bool GRAZ, bool CULL>
void launch_scan_gc(int which, int grid, cudaStream_t st, const FrameParams& P, int level) {
    constexpr size_t smem = sizeof(KernelSmem<2 * RP, CULL>);
    switch (which) {
        case 0: launch_kernel_smem(k_trace<RP, J, MINB, true, GRAZ, CULL>, smem, grid, st, P, level); break;
        case 1: launch_kernel_smem(k_trace<RP, J, MINB, false, GRAZ, CULL>, smem, grid, st, P, level); break;
        case 2: launch_kernel_smem(k_shadow<RP, J, MINB, false, GRAZ, CULL>, smem, grid, st, P, level); break;
        default: launch_kernel_smem(k_shadow<RP, J, MINB, true, GRAZ, CULL>, smem, grid, st, P, level); break;
    }
}
template <int RP, int J, int MINB>
void launch_scan(int which, int grid, cudaStream_t st, const FrameParams& P, int level, bool grazing_clause) {
    if (P.cull) {
        if (grazing_clause) launch_scan_gc<RP, J, MINB, true, true>(which, grid, st, P, level);
        else launch_scan_gc<RP, J, MINB, false, true>(which, grid, st, P, level);
    } else {
        if (grazing_clause) launch_scan_gc<RP, J, MINB, true, false>(which, grid, st, P, level);
        else launch_scan_gc<RP, J, MINB, false, false>(which, grid, st, P, level);
    }
}

// which: kScan*; the grid is one CTA per resident slot (persistent CTAs stride over ray chunks)
void dispatch_scan(const ScanConfig& c, int which, int num_sms, cudaStream_t st, const FrameParams& P, int level, bool grazing_clause) {
#define RT_X(RP, J, MINB) if (c.rp == RP && c.j == J && c.minb == MINB) return launch_scan<RP, J, MINB>(which, num_sms * MINB, st, P, level, grazing_clause);
    RT_SCAN_CONFIGS(RT_X)
#undef RT_X
}

template <int RP, int J, int MINB>
void launch_pencil(int which, int grid, cudaStream_t st, const FrameParams& P, int level) {
    constexpr size_t smem = sizeof(KernelSmem<2 * RP, false>);
    if (which == kScanPrimaryPencil) launch_kernel_smem(k_trace<RP, J, MINB, true, false, false, true>, smem, grid, st, P, level);
    else if (which == kScanBouncePencil) launch_kernel_smem(k_trace<RP, J, MINB, false, false, false, true>, smem, grid, st, P, level);
    else launch_kernel_smem(k_shadow<RP, J, MINB, false, false, false, true>, smem, grid, st, P, level);
}
void dispatch_pencil(const ScanConfig& c, int which, int num_sms, cudaStream_t st, const FrameParams& P, int level) {
#define RT_X(RP, J, MINB) if (c.rp == RP && c.j == J && c.minb == MINB) return launch_pencil<RP, J, MINB>(which, num_sms * MINB, st, P, level);
    RT_PENCIL_CONFIGS(RT_X)
#undef RT_X
}
bool pencil_config_exists(const ScanConfig& c) {
#define RT_X(RP, J, MINB) if (c.rp == RP && c.j == J && c.minb == MINB) return true;
    RT_PENCIL_CONFIGS(RT_X)
#undef RT_X
    return false;
}

bool scan_config_exists(const ScanConfig& c) {
#define RT_X(RP, J, MINB) if (c.rp == RP && c.j == J && c.minb == MINB) return true;
    RT_SCAN_CONFIGS(RT_X)
#undef RT_X
    return false;
}

void read_tuning_env() {
    if (const char* c = getenv("RT_B200_COSMIN")) {
        const float v = (float)atof(c);
        if (v >= 1e-6f && v <= 0.1f) g.cos_min = v;
    }
    if (const char* c = getenv("RT_B200_GRAZING")) g.allow_no_grazing = atoi(c) == 0;
    if (const char* c = getenv("RT_B200_CULL")) g.tile_culling = atoi(c) != 0;   // same as rt_set_option(RT_OPT_TILE_CULLING, ..)
    if (const char* c = getenv("RT_B200_PENCIL")) g.pencil = atoi(c) != 0;       // same as rt_set_option(RT_OPT_PENCIL, ..)
    if (const char* c = getenv("RT_B200_PENCIL_ANY")) g.pencil_any = atoi(c) != 0;   // same as rt_set_option(RT_OPT_PENCIL_ANY, ..)
    if (const char* c = getenv("RT_B200_PENCIL_REFLECT")) g.pencil_reflect = atoi(c) != 0;   // same as rt_set_option(RT_OPT_PENCIL_REFLECT, ..)
    if (const char* c = getenv("RT_B200_SMALL_TRACE")) g.small_trace = atoi(c) != 0;
    if (const char* c = getenv("RT_B200_PENCIL_THREAD")) g.pencil_thread = std::min(2, std::max(0, atoi(c)));   // same as rt_set_option(RT_OPT_PENCIL_THREAD, ..)
    if (const char* c = getenv("RT_B200_GRAPH")) g.graph_mode = atoi(c) < 0 ? -1 : (atoi(c) != 0);   // same as rt_set_option(RT_OPT_GRAPH, ..)
    if (const char* pe = getenv("RT_B200_PTUNE")) {
        ScanConfig c = g.pscan;
        if (sscanf(pe, "%d,%d,%d", &c.rp, &c.j, &c.minb) == 3 && pencil_config_exists(c)) g.pscan = c;
        else fprintf(stderr, "librt_b200: RT_B200_PTUNE=%s is not a compiled pencil configuration; keeping %d,%d,%d\n", pe, g.pscan.rp, g.pscan.j, g.pscan.minb);
    }
    const char* e = getenv("RT_B200_TUNE");
    if (!e) return;
    ScanConfig c = g.scan;
    if (sscanf(e, "%d,%d,%d", &c.rp, &c.j, &c.minb) == 3 && scan_config_exists(c)) g.scan = c;
    else fprintf(stderr, "librt_b200: RT_B200_TUNE=%s is not a compiled scan configuration; keeping %d,%d,%d\n", e, g.scan.rp, g.scan.j, g.scan.minb);
}

int check_ready() {
    if (g.devs.empty()) return fail(RT_ERR_STATE, "rt_init has not been called (or failed): there is no CPU fallback");
    return RT_OK;
}

float pow2_ceil(float v) {
    float m = 1.0f / 1024.0f;
    while (m < v) m *= 2.0f;
    return m;
}

// (Re)build the filter records when the magnitude bound M or the longest ray of the frame grows.
//
// Grazing clause.  A pair whose filter cosine is below cos_min must normally go to the exact path (the error bounds
// blow up there).  But the reference itself rejects a pair when |b| = |n.dir| < 1e-5 (raytracing.cpp:115) with the
// UNNORMALISED n = u x v and dir = dest - origin.  The filter's cosine is within 8u of the true one (3u FMA chain, u
// normal rounding, 4u rsqrt-normalised direction), so such a pair has |cos| < 1.05e-5; the reference's n is within
// 7u|u||v| of u x v (rounded edges and cross product), its dot product adds 3u|n||dir|, and |n| <= |u||v|:
//     |b_ref| < |dir| |u||v| (1.05e-5 + 7u + 3u) < 1.11e-5 |dir||u||v|.
// So if  max_triangles(|u||v|) * max_rays(|dir|) <= 0.85  every such pair is a certain miss in the reference and the
// clause can be compiled out (kernel variants with GRAZ = false) -- provided the face normals handed in are not longer
// than 1 (they bound the reflected / refracted directions in direction_bound()).  ("Always exact" triangles do not go
// through the filter at all: always_list.)  Fine meshes (the Balls stand-in, the 1 M-triangle sphere)
// qualify; scenes with large triangles (cube, ground quads) or far lights keep the clause.
bool clause_free_for(float dir_max) {
    return g.allow_no_grazing && g.cos_min <= 1.0e-5f && g.unit_normals && (double)g.max_uv * (double)dir_max <= 0.85;
}

// The records are (re)built for the request at hand, not for the largest one ever seen: one rt_trace call with a long
// ray or a far origin must not switch the clause-free kernels (and with them the pencil launches' premise) off for every
// later frame.  They are kept while they cover the request (M within a factor 4, every ray no longer than dir_built) AND
// decide the grazing clause the way this request would.  One host synchronisation per rebuild (n_always + scene box).
int build_records(RtDevice& d, float M, float dir_max) {
    dir_max = pow2_ceil(dir_max);   // coarse steps: a moving camera does not rebuild every frame
    const bool want_clause_free = clause_free_for(dir_max);
    if (d.rec && d.M_built >= M && d.M_built <= 4.0f * M && d.dir_built >= dir_max && d.no_grazing == want_clause_free) return RT_OK;
    CU(cudaSetDevice(d.device));
    const int npad = (d.ntiles + kPadTiles) * kTile;
    if (!d.n_always) CU(cudaMalloc(&d.n_always, sizeof(unsigned int)));
    CU(cudaMemsetAsync(d.n_always, 0, sizeof(unsigned int), d.stream));
    int rc = ensure(d.always_list, d.cap_always, (size_t)std::max(d.ntri, 1));
    if (rc) return rc;
    d.no_grazing = want_clause_free;
    d.M_built = 0.f;   // not valid until the read-back below has completed
    k_build_records<<<(npad + 127) / 128, 128, 0, d.stream>>>(d.triv, d.perm, npad, d.cls1 * kTile, d.cls2 * kTile, M,
                                                           d.no_grazing ? kBminNoGrazing : g.cos_min, d.rec, d.n_always, d.always_list);
    CU(cudaGetLastError());
    const int tiles_padded = d.ntiles + kPadTiles;
    k_build_tile_boxes<<<(tiles_padded + 127) / 128, 128, 0, d.stream>>>(d.triv, d.rec, tiles_padded, M, d.tile_box);
    CU(cudaGetLastError());
    const int nsuper = (tiles_padded + kSuper - 1) / kSuper;
    rc = ensure(d.super_box, d.cap_super, (size_t)nsuper * 2);
    if (rc) return rc;
    k_build_super_boxes<<<(nsuper + 127) / 128, 128, 0, d.stream>>>(d.tile_box, tiles_padded, nsuper, d.super_box);
    CU(cudaGetLastError());
    if (!d.scene_box) CU(cudaMalloc(&d.scene_box, 2 * sizeof(float4)));
    k_scene_box<<<1, 32, 0, d.stream>>>(d.super_box, nsuper, d.scene_box);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(d.h_small, d.scene_box, 2 * sizeof(float4), cudaMemcpyDeviceToHost, d.stream));
    CU(cudaMemcpyAsync(d.h_small + 8, d.n_always, sizeof(unsigned int), cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    unsigned int n_always = 0;
    memcpy(&n_always, d.h_small + 8, sizeof(n_always));
    d.n_always_host = (int)n_always;
    for (int k = 0; k < 3; ++k) { d.box_lo[k] = d.h_small[k]; d.box_hi[k] = d.h_small[4 + k]; }
    d.M_built = M;
    d.dir_built = dir_max;
    ++d.rec_gen;
    return RT_OK;
}

int ensure_chunk_state(RtDevice& d, size_t nsamples, bool want_prim, size_t prim_total) {
    CU(cudaSetDevice(d.device));
    if (nsamples > d.cap_samples) {
        void** ptrs[] = {(void**)&d.ray_o, (void**)&d.ray_d, (void**)&d.thr, (void**)&d.acc, (void**)&d.hit, (void**)&d.lit, (void**)&d.q_ray, (void**)&d.q_hit, (void**)&d.key};
        const size_t sizes[] = {16, 16, 16, 16, 16, 4, 4, 4, 8};
        for (int i = 0; i < 9; ++i) {
            if (*ptrs[i]) cudaFree(*ptrs[i]);
            *ptrs[i] = nullptr;
            CU(cudaMalloc(ptrs[i], sizes[i] * nsamples));
        }
        // keys are kKeyEmpty whenever no scan is in flight: set once here, restored by k_finish after every read
        CU(cudaMemsetAsync(d.key, 0xff, 8 * nsamples, d.stream));
        d.cap_samples = nsamples;
        d.key_dirty = false;
    }
    if (d.key_dirty) {
        CU(cudaMemsetAsync(d.key, 0xff, 8 * d.cap_samples, d.stream));
        d.key_dirty = false;
    }
    if (want_prim) { int rc = ensure(d.prim, d.cap_prim, prim_total); if (rc) return rc; }
    return RT_OK;
}

int ensure_counters(RtDevice& d, int slots) {
    if (slots <= d.counters_slots && d.counters) return RT_OK;
    if (d.counters) cudaFree(d.counters);
    d.counters = nullptr;
    CU(cudaMalloc(&d.counters, sizeof(uint32_t) * kCntWords * slots));
    d.counters_slots = slots;
    return RT_OK;
}

void fill_common(FrameParams& P, const RtDevice& d, const rt_params& rp, float eps_r, uint32_t* counters) {
    memset(&P, 0, sizeof(P));
    P.rec = d.rec; P.triv = d.triv; P.normal_mat = d.normal_mat; P.materials = d.materials; P.spheres = d.spheres;
    P.ntri = d.ntri; P.ntiles = d.ntiles; P.nspheres = d.nspheres; P.cls1 = d.cls1; P.cls2 = d.cls2;
    P.ray_o = d.ray_o; P.ray_d = d.ray_d; P.thr = d.thr; P.acc = d.acc; P.hit = d.hit; P.lit = d.lit;
    P.q_ray = d.q_ray; P.q_hit = d.q_hit; P.counters = counters; P.key = d.key;
    P.tile_box = d.tile_box; P.super_box = d.super_box;
    P.cull = (g.tile_culling && d.ntiles <= kCullMaxTiles) ? 1 : 0;   // beyond the bitmap's reach: brute force
    P.always_list = d.always_list; P.n_always = d.n_always_host;
    P.eps_r = eps_r;
    memcpy(P.camera, rp.camera, sizeof(P.camera));
    P.nlights = (int)rp.n_lights;
    memcpy(P.lights, rp.lights, sizeof(P.lights));
    P.features = rp.features;
    P.max_lvl = rp.max_lvl;
    P.light_sel = -1;
    P.mirror_sel = -1;
    P.n_mirrors = 0;
}


// Pencil launches of one frame on one device (rt_pencil.h).  cam: the primary rays; light[l]: the shadow rays of light l.
struct PencilPlan {
    bool cam = false, any_light = false;
    bool light[RT_MAX_LIGHTS] = {};
    bool no_premise = false;   // built without the clause-free proof (RT_OPT_PENCIL_ANY)
    PencilSetup cam_setup, light_setup[RT_MAX_LIGHTS];
    size_t slot_vec = 0;   // float4 per record slot
    // reflection pencils: mirror[g] serves the level-1 continuation rays of plane group g (record slot mirror_slot0 + g)
    int n_mirrors = 0;     // == number of plane groups when any of them qualifies (groups that do not get an impossible check)
    bool mirror[kMaxMirrors] = {};
    PencilSetup mirror_setup[kMaxMirrors];
    MirrorCheck mirror_check[kMaxMirrors];
    int mirror_slot0 = 0;
    // thread pencils: level-1 continuation rays of the primary hits on any other triangle (rt_tpencil.h)
    bool tp = false;
    TpSetup tp_setup;
};

// Thread pencils of this frame (buffers: ensure_tp).
void apply_tp(FrameParams& P, const RtDevice& d, const PencilPlan& plan) {
    P.tp_on = 1;
    P.trec = d.trec;
    P.tp = plan.tp_setup;
    P.tp_pool = d.tp_pool; P.q_tp = d.q_tp; P.tp_group_tri = d.tp_group_tri;
    P.tp_hist = d.tp_hist; P.tp_off = d.tp_off; P.tp_cursor = d.tp_cursor;
}
int ensure_tp(RtDevice& d) {
    if (d.cap_tp < d.cap_samples) {
        for (uint32_t** p : {&d.tp_pool, &d.q_tp, &d.tp_group_tri}) { if (*p) cudaFree(*p); *p = nullptr; CU(cudaMalloc(p, sizeof(uint32_t) * d.cap_samples)); }
        d.cap_tp = d.cap_samples;
    }
    const size_t nt = (size_t)std::max(d.ntri, 1);
    if (d.cap_tp_tri < nt) {
        for (uint32_t** p : {&d.tp_hist, &d.tp_off, &d.tp_cursor}) { if (*p) cudaFree(*p); *p = nullptr; CU(cudaMalloc(p, sizeof(uint32_t) * nt)); }
        d.cap_tp_tri = nt;
    }
    return RT_OK;
}

// Reflection pencils of this frame: what k_shade needs to route the level-1 continuation rays (level 0 only reads it).
void apply_mirrors(FrameParams& P, const RtDevice& d, const PencilPlan& plan) {
    P.n_mirrors = plan.n_mirrors;
    P.tri_group = d.tri_group;
    P.q_mirror = d.q_mirror;
    P.q_mirror_stride = (uint32_t)d.cap_samples;
    for (int k = 0; k < plan.n_mirrors; ++k) P.mirror[k] = plan.mirror_check[k];
}

void apply_pencil(FrameParams& P, const RtDevice& d, const PencilPlan& plan, int slot, const PencilSetup& S) {
    P.prec = d.prec + (size_t)slot * plan.slot_vec;
    P.pE[0] = S.Ef[0]; P.pE[1] = S.Ef[1]; P.pE[2] = S.Ef[2];
    P.p_lam_slack = S.lam_slack;
    memcpy(P.pF, S.F, sizeof(P.pF));
    P.p_wmax2 = S.w_max2;
}

// Decide which launches of this frame can use the pencil filter and build their records.  Conditions: brute force
// (no tile culling), the clause-free proof holds (pairs with |cos| < cos_min are certain misses in the reference), and
// the geometric launch conditions of pencil_camera_setup / pencil_light_setup.
int plan_pencil(RtDevice& d, const rt_params& rp, bool cull, PencilPlan& plan) {
    memset(static_cast<void*>(&plan), 0, sizeof(plan));   // (the plan's bytes are part of the graph key: no stack garbage in unused entries)
    new (&plan) PencilPlan;                               // default-initialisation: the member initialisers run, the rest stays zero
    // premise: the scene-level clause-free proof holds (pairs below cos_min are certain misses in the reference).  Without it
    // the pencil is used only under RT_OPT_PENCIL_ANY (experimental): near-plane triangles become "always candidate" records.
    const bool premise = d.no_grazing;
    if (!g.pencil || cull || d.ntri == 0 || (!premise && !g.pencil_any)) return RT_OK;
    // The pencil records depend on the generic records (rec_gen: scene, M, clause), the corner rays and the lights: an
    // unchanged frame set-up (a bench loop, a still camera) reuses them -- no build launches, no read-back.
    std::vector<uint8_t> key(sizeof(uint64_t) + sizeof(rp.corners) + sizeof(rp.lights) + 4 * sizeof(uint32_t));
    {
        uint8_t* k = key.data();
        memcpy(k, &d.rec_gen, sizeof(uint64_t)); k += sizeof(uint64_t);
        memcpy(k, rp.corners, sizeof(rp.corners)); k += sizeof(rp.corners);
        memcpy(k, rp.lights, sizeof(rp.lights)); k += sizeof(rp.lights);
        const uint32_t w[4] = {rp.n_lights, (rp.features & (RT_SHADOWS | RT_REFLECTION)) | (rp.max_lvl > 0 ? 1u << 16 : 0u),
                               (uint32_t)g.pencil_any | ((uint32_t)g.pencil_reflect << 1) | ((uint32_t)g.pencil_thread << 2), (uint32_t)g.any_transparent};
        memcpy(k, w, sizeof(w));
    }
    if (d.prec && key == d.plan_key && d.plan_blob.size() == sizeof(PencilPlan)) { memcpy(&plan, d.plan_blob.data(), sizeof(PencilPlan)); return RT_OK; }
    d.plan_key.clear();
    const int npad = (d.ntiles + kPadTiles) * kTile;
    plan.slot_vec = (size_t)npad * kRecVec;
    const bool shadows = (rp.features & RT_SHADOWS) && rp.n_lights > 0 && !g.any_transparent;
    // A pencil answers only for pairs with |cos| >= cos_g (>= the generic 1.05e-5): the clause-free proof of
    // build_records must hold at that threshold for the launch's rays, |b_ref| < |dir||u||v| (cos_g + 10u) < 1e-5.
    auto proof_holds = [&](const PencilSetup& S, double dir_max) { return !premise || (S.cos_g + 6e-7) * (double)g.max_uv * dir_max * 1.01 <= 0.95e-5; };
    plan.cam = pencil_camera_setup(rp.corners, (double)d.M_built, d.box_lo, d.box_hi, plan.cam_setup);
    if (plan.cam) {
        double dir_cam = 0.0;
        for (int c = 0; c < 4; ++c) {
            double l2 = 0.0;
            for (int k = 0; k < 3; ++k) { const double t = (double)rp.corners[c * 6 + 3 + k] - rp.corners[c * 6 + k]; l2 += t * t; }
            dir_cam = std::max(dir_cam, std::sqrt(l2));
        }
        plan.cam = proof_holds(plan.cam_setup, dir_cam);
    }
    if (shadows)
        for (uint32_t l = 0; l < rp.n_lights; ++l) {
            // shadow-ray length: |hit + 0.1 - light| for a hit inside the scene bound (as in direction_bound)
            double l2 = 0.0;
            for (int k = 0; k < 3; ++k) l2 += (double)rp.lights[l][k] * rp.lights[l][k];
            const double dir_l = std::sqrt(l2) + 1.7320508 * ((double)g.scene_extent + 0.1);
            plan.light[l] = pencil_light_setup(rp.lights[l], d.box_lo, d.box_hi, (double)d.M_built, plan.light_setup[l]) && proof_holds(plan.light_setup[l], dir_l);
            plan.any_light = plan.any_light || plan.light[l];
        }
    if (!plan.cam && !plan.any_light) return RT_OK;
    // reflection pencils: the level-1 continuation rays of primary hits on a plane group leave the mirror image of the eye
    plan.mirror_slot0 = 1 + (shadows ? (int)rp.n_lights : 0);
    if (plan.cam && g.pencil_reflect && (rp.features & RT_REFLECTION) && rp.max_lvl > 0 && !g.planes.empty()) {
        bool any = false;
        for (size_t k = 0; k < g.planes.size() && k < (size_t)kMaxMirrors; ++k) {
            // continuation rays are unit long (reflect_ray: dest = P + R, origin = P + 0.01 R)
            plan.mirror[k] = g.planes[k].count >= 2 &&
                             pencil_mirror_setup(plan.cam_setup, g.planes[k].n, g.planes[k].d, (double)d.M_built, d.box_lo, d.box_hi, plan.mirror_setup[k], plan.mirror_check[k]) &&
                             proof_holds(plan.mirror_setup[k], 1.05);
            if (!plan.mirror[k]) {   // not served: a check nothing passes
                memset(&plan.mirror_check[k], 0, sizeof(MirrorCheck));
                plan.mirror_check[k].w_max2 = -1.0f;
            }
            any = any || plan.mirror[k];
        }
        plan.n_mirrors = any ? (int)std::min(g.planes.size(), (size_t)kMaxMirrors) : 0;
    }
    // (not for launch-bound frames: grouping adds four launches and two memsets per chunk -- cube.obj 800x800 went 0.23 -> 0.36 ms)
    const double frame_tests = (double)rp.width * rp.height * rp.pixelfactor_x * rp.pixelfactor_y / (double)std::max(g.world, 1) * (double)std::max(d.ntri, 1);
    if (plan.cam && g.pencil_thread && (rp.features & RT_REFLECTION) && rp.max_lvl > 0 && (frame_tests > kGraphMaxTests || g.pencil_thread == 2))
        plan.tp = tp_setup(plan.cam_setup.E, plan.cam_setup.delta, (double)d.M_built, d.box_lo, d.box_hi, plan.tp_setup);
    int rc = ensure(d.prec, d.cap_prec, plan.slot_vec * (size_t)(plan.mirror_slot0 + plan.n_mirrors));
    if (rc) return rc;
    if (plan.tp) {
        rc = ensure(d.trec, d.cap_trec, (size_t)npad * kTpVec); if (rc) return rc;
        k_build_trec<<<(npad + 127) / 128, 128, 0, d.stream>>>(d.triv, d.rec, npad, d.cls1 * kTile, d.cls2 * kTile, d.M_built, plan.tp_setup, d.trec);
        CU(cudaGetLastError());
    }
    constexpr int kNearWords = 1 + RT_MAX_LIGHTS + kMaxMirrors;
    if (!d.n_near) CU(cudaMalloc(&d.n_near, sizeof(unsigned int) * kNearWords));
    if (!premise) CU(cudaMemsetAsync(d.n_near, 0, sizeof(unsigned int) * kNearWords, d.stream));
    const int grid = (npad + 127) / 128;
    if (plan.cam)
        k_build_pencil<<<grid, 128, 0, d.stream>>>(d.triv, d.rec, npad, d.cls1 * kTile, d.cls2 * kTile, d.M_built, plan.cam_setup, d.prec, premise ? 1 : 0, d.n_near);
    for (uint32_t l = 0; shadows && l < rp.n_lights; ++l)
        if (plan.light[l])
            k_build_pencil<<<grid, 128, 0, d.stream>>>(d.triv, d.rec, npad, d.cls1 * kTile, d.cls2 * kTile, d.M_built, plan.light_setup[l],
                                                       d.prec + (size_t)(1 + l) * plan.slot_vec, premise ? 1 : 0, d.n_near + 1 + l);
    for (int k = 0; k < plan.n_mirrors; ++k)
        if (plan.mirror[k])
            k_build_pencil<<<grid, 128, 0, d.stream>>>(d.triv, d.rec, npad, d.cls1 * kTile, d.cls2 * kTile, d.M_built, plan.mirror_setup[k],
                                                       d.prec + (size_t)(plan.mirror_slot0 + k) * plan.slot_vec, premise ? 1 : 0, d.n_near + 1 + RT_MAX_LIGHTS + k);
    CU(cudaGetLastError());
    if (!premise) {
        // too many "always candidate" records would turn the scan into an exact scan: such a launch keeps the generic kernels
        unsigned int h_near[kNearWords];
        CU(cudaMemcpyAsync(d.h_small + 16, d.n_near, sizeof(h_near), cudaMemcpyDeviceToHost, d.stream));
        CU(cudaStreamSynchronize(d.stream));
        memcpy(h_near, d.h_small + 16, sizeof(h_near));
        if (h_near[0] > kPencilMaxNear) plan.cam = false;
        plan.any_light = false;
        for (uint32_t l = 0; l < rp.n_lights; ++l) {
            if (h_near[1 + l] > kPencilMaxNear) plan.light[l] = false;
            plan.any_light = plan.any_light || plan.light[l];
        }
        bool any_mirror = false;
        for (int k = 0; k < plan.n_mirrors; ++k) {
            if (plan.mirror[k] && h_near[1 + RT_MAX_LIGHTS + k] > kPencilMaxNear) {
                plan.mirror[k] = false;
                memset(&plan.mirror_check[k], 0, sizeof(MirrorCheck));
                plan.mirror_check[k].w_max2 = -1.0f;
            }
            any_mirror = any_mirror || plan.mirror[k];
        }
        if (!any_mirror || !plan.cam) plan.n_mirrors = 0;
        plan.no_premise = plan.cam || plan.any_light;
    }
    d.plan_key = key;
    d.plan_blob.resize(sizeof(PencilPlan));
    memcpy(d.plan_blob.data(), &plan, sizeof(PencilPlan));
    return RT_OK;
}

// Wavefront for one chunk whose rays are either generated (primary) or already in ray_o/ray_d (trace API).
int run_wavefront(RtDevice& d, const FrameParams& P, float4* level0_hits = nullptr, const PencilPlan* plan = nullptr) {
    const bool shadows = (P.features & RT_SHADOWS) && P.nlights > 0;
    const bool bounces = (P.features & (RT_REFLECTION | RT_REFRACTION)) != 0;
    const int levels = bounces ? std::min(P.max_lvl + 1, kMaxLevels - 2) : 1;
    const int grid_small = d.num_sms * 4;
    for (int level = 0; level < levels; ++level) {
        if (level == 1 && plan && P.n_mirrors > 0 && !P.trace_api) {
            // the queues k_shade filled at level 0: one pencil scan + finish per plane group around the mirrored eye
            for (int gk = 0; gk < P.n_mirrors; ++gk) {
                if (!plan->mirror[gk]) continue;
                FrameParams Pm = P;
                apply_pencil(Pm, d, *plan, plan->mirror_slot0 + gk, plan->mirror_setup[gk]);
                Pm.mirror_sel = gk;
                {
                    LaunchTimer t(d, kKindTraceMirror);
                    dispatch_pencil(g.pscan, kScanBouncePencil, d.num_sms, d.stream, Pm, level);
                }
                {
                    LaunchTimer t(d, kKindShade);
                    k_finish<false><<<grid_small, 256, 0, d.stream>>>(Pm, level);
                }
            }
        }
        if (level == 1 && P.tp_on && !P.trace_api) {
            // thread pencils: group the pool k_shade filled at level 0 by reflector, scan the groups, finish them; whatever did not
            // fill a group has been appended to the ordinary queue, which the generic launch below serves
            {
                LaunchTimer t(d, kKindShade);
                k_tp_offsets<<<1, 1024, 0, d.stream>>>(P);
            }
            {
                LaunchTimer t(d, kKindShade);
                k_tp_scatter<<<grid_small, 256, 0, d.stream>>>(P);
            }
            FrameParams Pt = P;
            Pt.mirror_sel = kQueueTp;
            {
                LaunchTimer t(d, kKindTraceThread);
                launch_kernel_smem(k_trace_tp<4, 2>, sizeof(TpSmem), d.num_sms * 2, d.stream, Pt, level);
            }
            {
                LaunchTimer t(d, kKindShade);
                k_finish<false><<<grid_small, 256, 0, d.stream>>>(Pt, level);
            }
        }
        {
            LaunchTimer t(d, level == 0 ? kKindTracePrimary : kKindTrace);
            if (level == 0 && plan && plan->cam && !P.trace_api) {
                FrameParams Pp = P;
                apply_pencil(Pp, d, *plan, 0, plan->cam_setup);
                dispatch_pencil(g.pscan, kScanPrimaryPencil, d.num_sms, d.stream, Pp, level);
            } else {
                dispatch_scan(g.scan, level == 0 ? kScanPrimary : kScanBounce, d.num_sms, d.stream, P, level, !d.no_grazing);
            }
        }
        {
            LaunchTimer t(d, kKindShade);
            if (level == 0) k_finish<true><<<grid_small, 256, 0, d.stream>>>(P, 0);
            else k_finish<false><<<grid_small, 256, 0, d.stream>>>(P, level);
        }
        if (level == 0 && level0_hits) CU(cudaMemcpyAsync(level0_hits, d.hit, sizeof(float4) * P.nsamples, cudaMemcpyDeviceToDevice, d.stream));
        if (shadows && plan && plan->any_light) {
            // one launch per light: the pencil filter around the lights that qualify, the generic any-hit scan for the others
            for (int l = 0; l < P.nlights; ++l) {
                LaunchTimer t(d, kKindShadow);
                FrameParams Pl = P;
                Pl.light_sel = l;
                if (plan->light[l]) {
                    apply_pencil(Pl, d, *plan, 1 + l, plan->light_setup[l]);
                    dispatch_pencil(g.pscan, kScanShadowPencil, d.num_sms, d.stream, Pl, level);
                } else {
                    dispatch_scan(g.scan, kScanShadowAny, d.num_sms, d.stream, Pl, level, !d.no_grazing);
                }
            }
        } else if (shadows) {
            LaunchTimer t(d, kKindShadow);
            dispatch_scan(g.scan, g.any_transparent ? kScanShadowNearest : kScanShadowAny, d.num_sms, d.stream, P, level, !d.no_grazing);
        }
        {
            LaunchTimer t(d, kKindShade);
            k_shade<<<grid_small, 256, 0, d.stream>>>(P, level);
        }
    }
    CU(cudaGetLastError());
    return levels;
}

float magnitude_bound(const rt_params& rp, const float* extra, int n_extra) {
    float m = g.scene_extent + 0.5f;
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 3; ++k) m = std::max(m, std::fabs(rp.corners[c * 6 + k]));  // ray origins on the near plane
    for (int i = 0; i < n_extra; ++i) m = std::max(m, std::fabs(extra[i]));
    return pow2_ceil(m);
}

// Upper bound of |dest - origin| over every ray the wavefront can cast (see build_records): primary rays (bilinear
// blends of the corner rays, `frame`), shadow rays hit + 0.1 -> light, continuation rays (|R| = 1 for a reflection,
// |T| <= 1 + max(Ni, 1/Ni) for a refraction, minus the 0.01 offset), and the caller's rays for rt_trace.
float direction_bound(const rt_params& rp, bool frame, const float* origins, const float* dests, int n) {
    double m = 0.0;
    if (frame)
        for (int c = 0; c < 4; ++c) {
            double l2 = 0.0;
            for (int k = 0; k < 3; ++k) { const double t = (double)rp.corners[c * 6 + 3 + k] - rp.corners[c * 6 + k]; l2 += t * t; }
            m = std::max(m, std::sqrt(l2));
        }
    for (int i = 0; i < n; ++i) {
        double l2 = 0.0;
        for (int k = 0; k < 3; ++k) { const double t = (double)dests[3 * i + k] - origins[3 * i + k]; l2 += t * t; }
        m = std::max(m, std::sqrt(l2));
    }
    if ((rp.features & RT_SHADOWS) && rp.n_lights > 0) {
        const double reach = 1.7320508 * ((double)g.scene_extent + 0.1);   // |hit + 0.1| for a hit inside the scene box
        for (uint32_t i = 0; i < rp.n_lights; ++i) {
            double l2 = 0.0;
            for (int k = 0; k < 3; ++k) l2 += (double)rp.lights[i][k] * rp.lights[i][k];
            m = std::max(m, std::sqrt(l2) + reach);
        }
    }
    if (rp.features & RT_REFLECTION) m = std::max(m, 1.05);
    if (rp.features & RT_REFRACTION) m = std::max(m, 1.05 + (double)g.max_ni);
    if (!(m == m) || m > 1e30) m = 1e30;   // NaN / overflow: the clause stays
    return (float)(m * 1.01);
}

// Guard band of the distance tests (DESIGN.md "filter soundness"): the reference's r = a/b is within
// 24uM/|cos| of the true distance, its rounded hit point within 8uM of that, the filter's own r' within
// 10uM/|cos|; pairs with |cos| < cos_min never rely on it (they always go to the exact path).
float eps_r_for(float M) { return M * (48.0f * kU32 / g.cos_min + 128.0f * kU32); }

int validate_params(const rt_params* p, bool need_frame) {
    if (!p) return fail(RT_ERR_INVALID, "params is NULL");
    if (need_frame && (p->width == 0 || p->height == 0 || p->pixelfactor_x == 0 || p->pixelfactor_y == 0))
        return fail(RT_ERR_INVALID, "width/height/pixelfactor must be >= 1");
    if (p->n_lights > RT_MAX_LIGHTS) return fail(RT_ERR_INVALID, "at most %d lights", RT_MAX_LIGHTS);
    if (p->max_lvl < 0 || p->max_lvl > kMaxLevels - 3) return fail(RT_ERR_INVALID, "max_lvl must be in [0, %d]", kMaxLevels - 3);
    return RT_OK;
}

// Appends raw bytes to a cache key.
struct KeyWriter {
    std::vector<uint8_t>& v;
    template <class T> void put(const T& x) { const uint8_t* p = reinterpret_cast<const uint8_t*>(&x); v.insert(v.end(), p, p + sizeof(T)); }
    void put_bytes(const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); v.insert(v.end(), b, b + n); }
};

// Frames whose launches are short enough for the launch gaps to matter (C1: 44 launches for 0.6 ms of frame) are replayed
// from a captured CUDA graph.  Every scan launch has a fixed grid (persistent CTAs read their ray counts from device
// counters), so the whole wavefront -- all chunks, all levels, the resolve -- is capturable as it is.
constexpr int kSmallTraceRays = 32;        // rt_trace batches up to this size ...
constexpr double kSmallTraceTests = 1e5;   // ... and this many (ray, triangle) pairs per level take the single-launch path (measured: tools/trace_latency.py)

int render_enqueue_impl(const rt_params* rp) {
    int rc = check_ready();
    if (rc) return rc;
    if (!g.scene_ready) return fail(RT_ERR_STATE, "rt_render before rt_upload_scene");
    rc = validate_params(rp, true);
    if (rc) return rc;
    const uint32_t W = rp->width, H = rp->height, G = (uint32_t)g.world;
    const uint32_t spp = rp->pixelfactor_x * rp->pixelfactor_y;
    const uint32_t rows_per_rank = (H + G - 1) / G;
    const float M = magnitude_bound(*rp, nullptr, 0);
    const float eps_r = eps_r_for(M);
    const size_t row_samples = (size_t)W * spp;
    if (row_samples > kMaxChunkSamples * 4ull) return fail(RT_ERR_INVALID, "one row of samples (%zu) is too large", row_samples);
    uint32_t rows_per_chunk = (uint32_t)std::max<size_t>(1, kMaxChunkSamples / row_samples);
    if (rows_per_chunk >= 8) rows_per_chunk &= ~7u;   // whole 8x8-pixel blocks per chunk (slot_to_sample)

    for (RtDevice& d : g.devs) {
        CU(cudaSetDevice(d.device));
        (void)cudaGetLastError();   // see rt_trace
        d.kev_kind.clear();
        const uint32_t my_rows = (H > (uint32_t)d.rank) ? (H - d.rank + G - 1) / G : 0;
        const uint32_t nchunks = (my_rows + rows_per_chunk - 1) / rows_per_chunk;
        const size_t chunk_cap = (size_t)std::min(rows_per_chunk, std::max(my_rows, 1u)) * row_samples;
        rc = build_records(d, M, direction_bound(*rp, true, nullptr, nullptr, 0)); if (rc) return rc;
        PencilPlan plan;
        rc = plan_pencil(d, *rp, g.tile_culling && d.ntiles <= kCullMaxTiles, plan); if (rc) return rc;
        d.pencil_used = (plan.cam ? 2u : 0u) | (plan.any_light ? 4u : 0u) | (plan.no_premise ? 8u : 0u) | (plan.n_mirrors > 0 ? 32u : 0u) | (plan.tp ? 64u : 0u);
        rc = ensure_chunk_state(d, chunk_cap, rp->want_prim_id != 0, (size_t)rows_per_rank * row_samples); if (rc) return rc;
        if (plan.n_mirrors > 0) { rc = ensure(d.q_mirror, d.cap_q_mirror, (size_t)kMaxMirrors * d.cap_samples); if (rc) return rc; }
        if (plan.tp) { rc = ensure_tp(d); if (rc) return rc; }
        rc = ensure_counters(d, std::max(1u, nchunks)); if (rc) return rc;
        size_t need_local = (size_t)rows_per_rank * W * 3;
        rc = ensure(d.fb_local, d.cap_local, need_local); if (rc) return rc;
        if (G > 1) {
            rc = ensure(d.fb_gather, d.cap_gather, need_local * G); if (rc) return rc;
            rc = ensure(d.fb_final, d.cap_final, (size_t)W * H * 3); if (rc) return rc;
        }
        d.frame_chunks = (int)std::max(1u, nchunks);
        int levels_done = 0;
        // everything the device does for this rank's rows, in stream order (also what a graph capture records)
        auto body = [&]() -> int {
            CU(cudaMemsetAsync(d.counters, 0, sizeof(uint32_t) * kCntWords * std::max(1u, nchunks), d.stream));
            if (my_rows < rows_per_rank) CU(cudaMemsetAsync(d.fb_local, 0, need_local * sizeof(float), d.stream));
            if (rp->want_prim_id) CU(cudaMemsetAsync(d.prim, 0xff, sizeof(int32_t) * rows_per_rank * row_samples, d.stream));
            for (uint32_t c = 0; c < nchunks; ++c) {
                FrameParams P;
                fill_common(P, d, *rp, eps_r, d.counters + (size_t)c * kCntWords);
                memcpy(P.corners, rp->corners, sizeof(P.corners));
                P.divX = (float)(W * rp->pixelfactor_x - 1);   // main.cpp:360 (unsigned arithmetic, then float)
                P.divY = (float)(H * rp->pixelfactor_y - 1);   // main.cpp:361
                P.W = W; P.H = H; P.pfx = rp->pixelfactor_x; P.pfy = rp->pixelfactor_y;
                P.row0 = c * rows_per_chunk;
                P.nrows = std::min(rows_per_chunk, my_rows - P.row0);
                P.G = G; P.rank = (uint32_t)d.rank;
                P.nsamples = (uint32_t)(P.nrows * row_samples);
                P.tiles_x = (W + 7) / 8;
                P.nslots = P.tiles_x * ((P.nrows + 7) / 8) * 64u * spp;
                P.sample_base = (unsigned long long)P.row0 * row_samples;
                P.prim_out = rp->want_prim_id ? d.prim : nullptr;
                if (plan.n_mirrors > 0) apply_mirrors(P, d, plan);
                if (plan.tp) {
                    apply_tp(P, d, plan);
                    CU(cudaMemsetAsync(d.tp_hist, 0, sizeof(uint32_t) * (size_t)std::max(d.ntri, 1), d.stream));
                    CU(cudaMemsetAsync(d.tp_cursor, 0, sizeof(uint32_t) * (size_t)std::max(d.ntri, 1), d.stream));
                }
                int levels = run_wavefront(d, P, nullptr, &plan);
                if (levels < 0) return levels;
                levels_done = levels;
                {
                    LaunchTimer t(d, kKindResolve);
                    k_resolve<<<d.num_sms * 4, 256, 0, d.stream>>>(P, d.fb_local);
                }
                CU(cudaGetLastError());
            }
            return RT_OK;
        };
        const double tests = (double)my_rows * (double)row_samples * (double)std::max(d.ntri, 1);
        const bool use_graph = g.graph_mode > 0 || (g.graph_mode < 0 && tests <= kGraphMaxTests && !getenv("RT_B200_LAUNCHLOG"));
        if (!use_graph) {
            CU(cudaEventRecord(d.ev_phase[0], d.stream));
            rc = body(); if (rc) return rc;
            CU(cudaEventRecord(d.ev_phase[1], d.stream));
        } else {
            // the captured launches take their arguments by value: the key holds everything they are derived from
            std::vector<uint8_t> key;
            KeyWriter kw{key};
            kw.put(*rp); kw.put(d.rec_gen); kw.put_bytes(&plan, sizeof(plan));
            const void* ptrs[] = {d.rec, d.triv, d.normal_mat, d.materials, d.spheres, d.prec, d.tile_box, d.super_box, d.always_list, d.ray_o, d.ray_d,
                                  d.thr, d.acc, d.hit, d.lit, d.q_ray, d.q_hit, d.key, d.counters, d.prim, d.fb_local, d.q_mirror, d.tri_group, d.trec, d.tp_pool, d.q_tp, d.tp_group_tri, d.tp_hist, d.tp_off, d.tp_cursor};
            kw.put(ptrs);
            const int cfg[] = {g.scan.rp, g.scan.j, g.scan.minb, g.pscan.rp, g.pscan.j, g.pscan.minb, (int)g.tile_culling, (int)G, d.rank, d.num_sms, (int)g.any_transparent};
            kw.put(cfg);
            if (!d.frame_graph || key != d.graph_key) {
                if (d.frame_graph) { cudaGraphExecDestroy(d.frame_graph); d.frame_graph = nullptr; }
                d.graph_key.clear();
                CU(cudaStreamBeginCapture(d.stream, cudaStreamCaptureModeThreadLocal));
                d.capturing = true;
                rc = body();
                d.capturing = false;
                cudaGraph_t graph = nullptr;
                const cudaError_t e = cudaStreamEndCapture(d.stream, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (e != cudaSuccess) return fail(RT_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
                const cudaError_t e2 = cudaGraphInstantiate(&d.frame_graph, graph, 0);
                cudaGraphDestroy(graph);
                if (e2 != cudaSuccess) { d.frame_graph = nullptr; return fail(RT_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e2)); }
                d.graph_key = key;
                d.launches_in_graph = (uint32_t)d.kev_kind.size();
                d.levels_in_graph = levels_done;
            }
            d.kev_kind.assign(d.launches_in_graph, -1);   // counted, not timed individually
            levels_done = d.levels_in_graph;
            CU(cudaEventRecord(d.ev_phase[0], d.stream));
            CU(cudaGraphLaunch(d.frame_graph, d.stream));
            CU(cudaEventRecord(d.ev_phase[1], d.stream));
        }
        g.stats.n_levels = (uint32_t)levels_done;
        d.used_graph = use_graph;
    }
    // one all-gather per frame (rank-major slabs), then de-interleave rows
    if (G > 1 && !g.devs[0].comm) {
        // detached rank (rt_init_rank without an ncclUniqueId): no exchange, own rows only
        RtDevice& d = g.devs[0];
        {
            LaunchTimer t(d, kKindGather);
            k_place_rows<<<d.num_sms * 4, 256, 0, d.stream>>>(d.fb_local, d.fb_final, W, H, G, (uint32_t)d.rank);
        }
        CU(cudaGetLastError());
    } else if (G > 1) {
        const size_t count = (size_t)rows_per_rank * W * 3;
        if (g.single_process) NC(g.nccl.GroupStart());
        for (RtDevice& d : g.devs) {
            CU(cudaSetDevice(d.device));
            NC(g.nccl.AllGather(d.fb_local, d.fb_gather, count, kNcclFloat, d.comm, d.stream));
        }
        if (g.single_process) NC(g.nccl.GroupEnd());
        for (RtDevice& d : g.devs) {
            CU(cudaSetDevice(d.device));
            {
                LaunchTimer t(d, kKindGather);
                k_deinterleave<<<d.num_sms * 4, 256, 0, d.stream>>>(d.fb_gather, d.fb_final, W, H, G, rows_per_rank);
            }
            CU(cudaGetLastError());
        }
    }
    for (RtDevice& d : g.devs) {
        CU(cudaSetDevice(d.device));
        CU(cudaEventRecord(d.ev_phase[2], d.stream));
    }
    g.last = *rp;
    g.rows_per_rank = rows_per_rank;
    g.frame_ready = true;
    g.stats_ready = true;
    g.stats_of_trace = false;
    return RT_OK;
}

// The frame state is invalid while a frame is being enqueued; a call that fails half way may leave a scan without its
// k_finish (the merge keys are only reset there) and buffers reallocated: nothing of the old frame can be downloaded
// afterwards, and the keys are re-initialised before their next use.
int render_enqueue(const rt_params* rp) {
    g.frame_ready = g.stats_ready = false;
    const int rc = render_enqueue_impl(rp);
    if (rc != RT_OK)
        for (RtDevice& d : g.devs) { d.key_dirty = true; d.capturing = false; }
    return rc;
}

int sync_all() {
    for (RtDevice& d : g.devs) {
        CU(cudaSetDevice(d.device));
        CU(cudaStreamSynchronize(d.stream));
    }
    return RT_OK;
}

// Gather ray counters of the last frame from device 0..n (this process's share).
int collect_stats() {
    rt_stats& st = g.stats;
    const uint32_t levels = st.n_levels;
    st = rt_stats();
    st.n_levels = levels;
    st.n_gpus = (uint32_t)g.world;
    st.rank = g.devs.empty() ? 0 : (uint32_t)g.devs[0].rank;
    const rt_params& rp = g.stats_of_trace ? g.last_trace : g.last;
    const bool shadows = (rp.features & RT_SHADOWS) && rp.n_lights > 0;
    for (RtDevice& d : g.devs) {
        if (g.stats_of_trace && &d != &g.devs[0]) break;   // rt_trace runs on the first device only
        CU(cudaSetDevice(d.device));
        st.n_triangles = (uint32_t)d.ntri;
        g.host_counters.resize((size_t)kCntWords * d.counters_slots);
        CU(cudaMemcpy(g.host_counters.data(), d.counters, sizeof(uint32_t) * g.host_counters.size(), cudaMemcpyDeviceToHost));
        const uint32_t H = rp.height, G = (uint32_t)g.world;
        const uint32_t my_rows = (H > (uint32_t)d.rank) ? (H - d.rank + G - 1) / G : 0;
        st.primary_rays += g.stats_of_trace ? (uint64_t)g.last_trace_n : (uint64_t)my_rows * rp.width * rp.pixelfactor_x * rp.pixelfactor_y;
        for (int c = 0; c < d.frame_chunks; ++c) {
            const uint32_t* cw = &g.host_counters[(size_t)c * kCntWords];
            for (int l = 0; l < kMaxLevels; ++l) {
                if (shadows) st.shadow_rays += (uint64_t)cw[kCntHit + l] * rp.n_lights;
                if (l > 0) st.bounce_rays += cw[kCntRay + l];
            }
            for (int k = 0; k < kMaxMirrors; ++k) { st.bounce_rays += cw[kCntMirror + k]; st.mirror_rays += cw[kCntMirror + k]; }   // level-1 rays served by a mirror pencil
            st.bounce_rays += (uint64_t)cw[kCntTpGroups] * kTpR; st.thread_pencil_rays += (uint64_t)cw[kCntTpGroups] * kTpR;       // ... by thread pencils
            st.exact_evals += (uint64_t)cw[kCntExact] | ((uint64_t)cw[kCntExact + 1] << 32);
        }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, d.ev_phase[0], d.ev_phase[2]) == cudaSuccess) st.ms_total = std::max(st.ms_total, ms);
        float by_kind[kNumKinds] = {};
        for (size_t i = 0; i < d.kev_kind.size(); ++i)
            if (d.kev_kind[i] >= 0 && cudaEventElapsedTime(&ms, d.kev[2 * i], d.kev[2 * i + 1]) == cudaSuccess) by_kind[d.kev_kind[i]] += ms;
        if (getenv("RT_B200_LAUNCHLOG")) {   // per-launch device times of the frame, in launch order (diagnostics)
            static const char* names[kNumKinds] = {"trace", "shadow", "shade/finish", "resolve", "gather", "trace_primary", "trace_mirror", "trace_thread"};
            for (size_t i = 0; i < d.kev_kind.size(); ++i)
                if (d.kev_kind[i] >= 0 && cudaEventElapsedTime(&ms, d.kev[2 * i], d.kev[2 * i + 1]) == cudaSuccess && ms > 0.05f)
                    fprintf(stderr, "librt_b200: launch %3zu %-14s %9.3f ms\n", i, names[d.kev_kind[i]], ms);
            for (int c = 0; c < d.frame_chunks; ++c) {
                const uint32_t* cw = &g.host_counters[(size_t)c * kCntWords];
                for (int l = 0; l < 6; ++l) fprintf(stderr, "librt_b200: chunk %d level %d: %u rays in, %u hits\n", c, l, l == 0 ? 0u : cw[kCntRay + l], cw[kCntHit + l]);
            }
        }
        st.ms_trace = std::max(st.ms_trace, by_kind[kKindTrace] + by_kind[kKindTracePrimary] + by_kind[kKindTraceMirror] + by_kind[kKindTraceThread]);
        st.ms_trace_thread = std::max(st.ms_trace_thread, by_kind[kKindTraceThread]);
        st.ms_trace_mirror = std::max(st.ms_trace_mirror, by_kind[kKindTraceMirror]);
        st.ms_trace_primary = std::max(st.ms_trace_primary, by_kind[kKindTracePrimary]);
        st.ms_shadow = std::max(st.ms_shadow, by_kind[kKindShadow]);
        st.ms_shade = std::max(st.ms_shade, by_kind[kKindShade]);
        st.ms_resolve = std::max(st.ms_resolve, by_kind[kKindResolve]);
        if (cudaEventElapsedTime(&ms, d.ev_phase[1], d.ev_phase[2]) == cudaSuccess) st.ms_gather = std::max(st.ms_gather, ms);
        st.n_launches = std::max(st.n_launches, (uint32_t)d.kev_kind.size());
        st.variant |= (d.no_grazing ? 1u : 0u) | d.pencil_used | (d.used_graph ? 16u : 0u);   // pencil_used carries bit 5 (32) for mirror pencils
    }
    st.tri_tests = (st.primary_rays + st.shadow_rays + st.bounce_rays) * (uint64_t)st.n_triangles;
    // cudaEventElapsedTime on events a call never recorded (rt_trace records no frame phases) fails harmlessly above, but
    // the runtime keeps the code as its "last error": drop it, or the next cudaGetLastError() after a launch reports it
    (void)cudaGetLastError();
    return RT_OK;
}

}  // namespace

extern "C" {

const char* rt_last_error(void) { return g.error.c_str(); }

// Frees every device resource; the options (rt_set_option) survive, so an option set before rt_init -- which starts by
// releasing whatever an earlier rt_init left -- is not lost.
static void release_all() {
    for (RtDevice& d : g.devs) destroy_device(d);
    g.devs.clear();
    g.world = 0;
    g.scene_ready = g.frame_ready = g.stats_ready = false;
    g_stage_triv.release(); g_stage_nm.release(); g_stage_sph.release(); g_stage_perm.release(); g_stage_mat.release();
    g_stage_rays.release(); g_stage_out.release(); g_stage_group.release();
}

void rt_shutdown(void) {
    release_all();
    g.tile_culling = false;
    g.pencil = true;
    g.pencil_any = true;
    g.pencil_reflect = true;
    g.pencil_thread = 1;
    g.small_trace = true;
    g.graph_mode = -1;
}

int rt_init(int n_gpus) {
    release_all();
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(RT_ERR_NO_DEVICE, "no CUDA device (%s); librt_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count == 0");
    if (n_gpus < 1 || n_gpus > count) return fail(RT_ERR_INVALID, "n_gpus=%d but %d device(s) are visible", n_gpus, count);
    read_tuning_env();
    g.devs.resize(n_gpus);
    g.world = n_gpus;
    g.single_process = true;
    for (int i = 0; i < n_gpus; ++i) {
        int rc = create_device(g.devs[i], i, i);
        if (rc) { release_all(); return rc; }
    }
    if (n_gpus > 1) {
        if (!g.nccl.load()) { release_all(); return fail(RT_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror()); }
        std::vector<ncclComm_t> comms(n_gpus);
        std::vector<int> ids(n_gpus);
        for (int i = 0; i < n_gpus; ++i) ids[i] = i;
        ncclResult_t r = g.nccl.CommInitAll(comms.data(), n_gpus, ids.data());
        if (r != 0) { release_all(); return fail(RT_ERR_NCCL, "ncclCommInitAll failed: %s", g.nccl.GetErrorString(r)); }
        for (int i = 0; i < n_gpus; ++i) g.devs[i].comm = comms[i];
    }
    return RT_OK;
}

int rt_nccl_unique_id(void* out, size_t cap, size_t* bytes) {
    if (!g.nccl.load()) return fail(RT_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    if (cap < sizeof(ncclUniqueId)) return fail(RT_ERR_INVALID, "need %zu bytes for an ncclUniqueId", sizeof(ncclUniqueId));
    ncclUniqueId id;
    NC(g.nccl.GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    if (bytes) *bytes = sizeof(id);
    return RT_OK;
}

int rt_init_rank(int device, int rank, int world, const void* nccl_id, size_t nccl_id_bytes) {
    release_all();
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(RT_ERR_NO_DEVICE, "no CUDA device (%s); librt_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count == 0");
    if (device < 0 || device >= count || world < 1 || rank < 0 || rank >= world) return fail(RT_ERR_INVALID, "bad device/rank/world %d/%d/%d", device, rank, world);
    read_tuning_env();
    g.devs.resize(1);
    g.world = world;
    g.single_process = false;
    int rc = create_device(g.devs[0], device, rank);
    if (rc) { release_all(); return rc; }
    if (world > 1 && nccl_id) {
        if (!g.nccl.load()) { release_all(); return fail(RT_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror()); }
        if (nccl_id_bytes < sizeof(ncclUniqueId)) { release_all(); return fail(RT_ERR_INVALID, "ncclUniqueId needs %zu bytes", sizeof(ncclUniqueId)); }
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof(id));
        ncclResult_t r = g.nccl.CommInitRank(&g.devs[0].comm, world, id, rank);
        if (r != 0) { release_all(); return fail(RT_ERR_NCCL, "ncclCommInitRank failed: %s", g.nccl.GetErrorString(r)); }
    }
    return RT_OK;
}

int rt_upload_scene(const rt_scene* sc) {
    int rc = check_ready();
    if (rc) return rc;
    if (!sc || (sc->n_triangles && (!sc->v0 || !sc->v1 || !sc->v2 || !sc->normal || !sc->tri_material)) || !sc->materials || sc->n_materials == 0)
        return fail(RT_ERR_INVALID, "incomplete rt_scene");
    if (sc->n_spheres && !sc->spheres) return fail(RT_ERR_INVALID, "n_spheres > 0 but spheres is NULL");
    const uint32_t n = sc->n_triangles;
    for (uint32_t i = 0; i < sc->n_spheres; ++i)
        if (sc->spheres[i].material >= sc->n_materials) return fail(RT_ERR_INVALID, "sphere %u uses material %u of %u", i, sc->spheres[i].material, sc->n_materials);
    // the staging buffers are reused: the copies of the previous upload must have left them
    for (RtDevice& d : g.devs) {
        CU(cudaSetDevice(d.device));
        CU(cudaEventSynchronize(d.ev_stage));
    }

    // ONE pass over the triangles: validation, packing (exact corners: 3 float4 per triangle; normal + material), the
    // scene extent, the dominant axis of each plane normal (rt_kernels.cuh: 2-D projected filter) and the inputs of the
    // clause-free proof (build_records): max |u||v| and "no face normal longer than 1".  Float arithmetic is enough for
    // all of these: the class only has to be SOME axis with a non-zero normal component (checked again, in double, when
    // the record is built), max_uv carries a 1e-4 margin, and anything non-finite gives up the proof.
    PinnedStage<float4>&triv = g_stage_triv, &nm = g_stage_nm;
    triv.resize((size_t)3 * std::max(n, 1u));
    nm.resize(std::max(n, 1u));
    if (!triv.data() || !nm.data()) return fail(RT_ERR_CUDA, "out of host memory for the scene staging buffers");
    static std::vector<uint8_t> cls;
    cls.resize(n);
    // plane groups (reflection pencils): heavy hitters among the quantised plane equations, one direct-mapped table pass
    struct PlaneSlot { uint64_t key; uint32_t count; uint32_t first; };
    constexpr uint32_t kPlaneSlots = 4096;
    static std::vector<PlaneSlot> ptab;
    static std::vector<uint64_t> pkey;
    pkey.resize(n);
    // The pass runs on up to 4 host threads over contiguous triangle ranges (it is the fixed cost of every end-to-end step:
    // 1.3 ms single-threaded for the 44,672-triangle headline scene); every thread keeps its own bounds, class counts and plane
    // table, merged below.
    struct Part {
        size_t cnt[4] = {0, 0, 0, 0};
        float extent = 0.f;
        double max_uuvv = 0.0;
        bool unit_normals = true;
        long long bad_tri = -1;
        uint32_t bad_mat = 0;
        std::vector<PlaneSlot> ptab;
    };
    // (one process per GPU: the ranks of a box share its cores -- measured at 8 ranks on 32 cores, 4 threads per rank made the
    // slowest rank's upload slower, 1.8 against 1.5 ms; so the threads are rationed by the number of ranks)
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned share = g.single_process ? hw : std::max(1u, hw / (4u * (unsigned)std::max(g.world, 1)));
    const int nthreads = n >= 16384 ? (int)std::min(4u, share) : 1;
    static std::vector<Part> parts;
    parts.assign((size_t)nthreads, Part());
    for (Part& P : parts) P.ptab.assign(kPlaneSlots, PlaneSlot{0, 0, 0});
    auto work = [&](int t) {
        Part& P = parts[(size_t)t];
        const uint32_t i0 = (uint32_t)((uint64_t)n * t / nthreads), i1 = (uint32_t)((uint64_t)n * (t + 1) / nthreads);
        for (uint32_t i = i0; i < i1; ++i) {
            const uint32_t m = sc->tri_material[i];
            if (m >= sc->n_materials) { if (P.bad_tri < 0) { P.bad_tri = (long long)i; P.bad_mat = m; } continue; }
            const float *A = sc->v0 + 4 * i, *B = sc->v1 + 4 * i, *C = sc->v2 + 4 * i, *N = sc->normal + 4 * i;
            triv[3 * i] = make_float4(A[0], A[1], A[2], 0.f);
            triv[3 * i + 1] = make_float4(B[0], B[1], B[2], 0.f);
            triv[3 * i + 2] = make_float4(C[0], C[1], C[2], 0.f);
            float4 v = make_float4(N[0], N[1], N[2], 0.f);
            memcpy(&v.w, &m, 4);
            nm[i] = v;
            float e = std::fmax(std::fmax(std::fabs(A[0]), std::fabs(A[1])), std::fabs(A[2]));
            e = std::fmax(e, std::fmax(std::fmax(std::fabs(B[0]), std::fabs(B[1])), std::fabs(B[2])));
            e = std::fmax(e, std::fmax(std::fmax(std::fabs(C[0]), std::fabs(C[1])), std::fabs(C[2])));   // fmax drops NaN operands
            if (e <= FLT_MAX) P.extent = std::max(P.extent, e);
            else  // an infinite coordinate: the finite ones still count
                for (int k = 0; k < 3; ++k) { for (const float* p3 : {A, B, C}) if (std::isfinite(p3[k])) P.extent = std::max(P.extent, std::fabs(p3[k])); }
            const float ux = B[0] - A[0], uy = B[1] - A[1], uz = B[2] - A[2];
            const float vx = C[0] - A[0], vy = C[1] - A[1], vz = C[2] - A[2];
            const float nx = std::fabs(uy * vz - uz * vy), ny = std::fabs(uz * vx - ux * vz), nz = std::fabs(ux * vy - uy * vx);
            int w = 0;
            if (ny > nx) w = 1;
            if (nz > (w == 1 ? ny : nx)) w = 2;   // NaN compares false: non-finite triangles land in class 0 (and are "always exact")
            cls[i] = (uint8_t)w;
            ++P.cnt[w];
            {   // quantised plane (unit normal with a canonical sign, offset): coplanar triangles share the key (up to bin edges --
                // a split group only serves fewer rays: every ray is checked against its pencil when it is routed, k_shade)
                const float cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
                const float l2n = cx * cx + cy * cy + cz * cz;
                uint64_t key = 0;
                if (l2n > 0.f && l2n < 1e30f) {
                    float inv = 1.0f / std::sqrt(l2n);
                    if (cx < 0.f || (cx == 0.f && (cy < 0.f || (cy == 0.f && cz < 0.f)))) inv = -inv;
                    const float qx = cx * inv, qy = cy * inv, qz = cz * inv, qd = qx * A[0] + qy * A[1] + qz * A[2];
                    // round to nearest by truncating a positive number (components are in [-1, 1]; the offset is clamped)
                    const int64_t ix = (int64_t)(int)(qx * 16384.f + 16384.5f), iy = (int64_t)(int)(qy * 16384.f + 16384.5f), iz = (int64_t)(int)(qz * 16384.f + 16384.5f);
                    const float qdc = qd * 1024.f;
                    const int64_t id = (int64_t)(int)((qdc < -2.0e6f ? -2.0e6f : (qdc > 2.0e6f ? 2.0e6f : qdc)) + 2097152.5f);
                    key = ((uint64_t)ix << 48) ^ ((uint64_t)iy << 32) ^ ((uint64_t)iz << 16) ^ ((uint64_t)id * 0x9E3779B97F4A7C15ull) | 1ull;
                }
                pkey[i] = key;
                if (key) {
                    PlaneSlot& ps = P.ptab[(uint32_t)((key * 0xD6E8FEB86659FD93ull) >> 52)];
                    if (ps.key == key) ++ps.count;
                    else if (ps.count <= 1) { ps.key = key; ps.count = 1; ps.first = i; }
                    else --ps.count;   // Misra-Gries style: a heavy plane keeps its slot
                }
            }
            const double p2 = (double)(ux * ux + uy * uy + uz * uz) * (double)(vx * vx + vy * vy + vz * vz);
            P.max_uuvv = (p2 == p2) ? std::max(P.max_uuvv, p2) : INFINITY;   // NaN vertices: never claim the bound
            const float l2 = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
            if (!(l2 <= 1.00001f)) P.unit_normals = false;
        }
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t) {
            try { pool.emplace_back(work, t); }
            catch (...) { work(t); }   // no thread to be had (resource limits): do that range here -- nothing is thrown across the C ABI
        }
        work(0);
        for (std::thread& th : pool) th.join();
    }
    size_t cnt[4] = {0, 0, 0, 0};
    float extent = 0.f;
    double max_uuvv = 0.0;
    bool unit_normals = true;
    ptab.assign(kPlaneSlots, PlaneSlot{0, 0, 0});
    for (const Part& P : parts) {
        if (P.bad_tri >= 0) return fail(RT_ERR_INVALID, "triangle %lld uses material %u of %u", P.bad_tri, P.bad_mat, sc->n_materials);
        for (int c = 0; c < 4; ++c) cnt[c] += P.cnt[c];
        extent = std::max(extent, P.extent);
        max_uuvv = (P.max_uuvv == P.max_uuvv) ? std::max(max_uuvv, P.max_uuvv) : INFINITY;
        unit_normals = unit_normals && P.unit_normals;
        for (uint32_t k = 0; k < kPlaneSlots; ++k) {   // same hash, same slot: equal keys add up, otherwise the heavier plane stays
            const PlaneSlot& ps = P.ptab[k];
            PlaneSlot& out = ptab[k];
            if (!ps.key || !ps.count) continue;
            if (out.key == ps.key) out.count += ps.count;
            else if (ps.count > out.count) out = ps;
        }
    }
    g.max_uv = (float)std::min(std::sqrt(max_uuvv) * 1.0001, 1e30);
    g.unit_normals = unit_normals;
    // the (at most kMaxMirrors) heaviest planes become groups; exact membership counts in a second, cheap pass
    g.planes.clear();
    PinnedStage<uint8_t>& grp = g_stage_group;
    grp.resize(std::max(n, 1u));
    if (!grp.data()) return fail(RT_ERR_CUDA, "out of host memory for the scene staging buffers");
    {
        uint64_t top_key[kMaxMirrors]; uint32_t top_cnt[kMaxMirrors], top_first[kMaxMirrors]; int ntop = 0;
        // a group must be worth a set of pencil records and two launches per frame: at least 2 triangles and 0.1 % of the scene
        // (the two triangles of a cube face qualify; the planar quads of a finely tessellated sphere do not)
        const uint32_t min_count = std::max(2u, n / 1000u);
        for (const PlaneSlot& ps : ptab) {
            if (ps.count + 16 < min_count || ps.count < 2 || !ps.key) continue;   // (the table's counts are lower bounds: exact counts below)
            int at = ntop < kMaxMirrors ? ntop++ : -1;
            if (at < 0) { int w = 0; for (int k = 1; k < kMaxMirrors; ++k) if (top_cnt[k] < top_cnt[w]) w = k; if (top_cnt[w] < ps.count) at = w; }
            if (at >= 0) { top_key[at] = ps.key; top_cnt[at] = ps.count; top_first[at] = ps.first; }
        }
        uint32_t members[kMaxMirrors] = {};
        for (uint32_t i = 0; i < n; ++i) {
            uint8_t gi = kNoGroup;
            for (int k = 0; k < ntop; ++k) if (pkey[i] == top_key[k]) { gi = (uint8_t)k; ++members[k]; break; }
            grp[i] = gi;
        }
        for (int k = 0; k < ntop; ++k) {   // the plane of the group: its first member's, in double
            const uint32_t i = top_first[k];
            const float *A = sc->v0 + 4 * i, *B = sc->v1 + 4 * i, *C = sc->v2 + 4 * i;
            const double u[3] = {(double)B[0] - A[0], (double)B[1] - A[1], (double)B[2] - A[2]}, v[3] = {(double)C[0] - A[0], (double)C[1] - A[1], (double)C[2] - A[2]};
            double nn[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
            const double l = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
            Global::PlaneGroup pg;
            for (int a = 0; a < 3; ++a) pg.n[a] = l > 0.0 ? nn[a] / l : 0.0;
            pg.d = pg.n[0] * A[0] + pg.n[1] * A[1] + pg.n[2] * A[2];
            pg.count = members[k] >= min_count ? members[k] : 0;   // count 0: never served (plan_pencil)
            g.planes.push_back(pg);   // a degenerate first member (l == 0) fails pencil_mirror_setup later: the group is simply not served
        }
    }
    // group the triangles by class; stable inside a class, every class padded to whole tiles
    PinnedStage<uint32_t>& perm = g_stage_perm;
    int cls_tiles[3] = {0, 0, 0};
    {
        size_t start[4], total = 0;
        int tiles4[4];
        for (int c = 0; c < 4; ++c) { start[c] = total; tiles4[c] = (int)((cnt[c] + kTile - 1) / kTile); total += (size_t)tiles4[c] * kTile; }
        cls_tiles[0] = tiles4[0]; cls_tiles[1] = tiles4[1]; cls_tiles[2] = tiles4[2] + tiles4[3];   // class 3 rides on the W = z path
        if (total == 0) { cls_tiles[0] = 1; total = kTile; }   // an empty scene still has one (padding) tile
        perm.assign(total + (size_t)kPadTiles * kTile, kNoTriangle);
        if (!perm.data()) return fail(RT_ERR_CUDA, "out of host memory for the scene staging buffers");
        size_t fill[4] = {start[0], start[1], start[2], start[3]};
        for (uint32_t i = 0; i < n; ++i) perm[fill[cls[i]]++] = i;
        if (g.tile_culling && n > 0) {
            // tile culling wants spatially compact tiles: inside each class, order the triangles along a Morton curve
            // of their centroids (the scan order is free: ties are broken by triangle id, results merge through keys)
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            std::vector<float> cen((size_t)3 * n);
            for (uint32_t i = 0; i < n; ++i)
                for (int a = 0; a < 3; ++a) {
                    const float c = (sc->v0[4 * i + a] + sc->v1[4 * i + a] + sc->v2[4 * i + a]) * (1.0f / 3.0f);
                    cen[3 * (size_t)i + a] = c;
                    if (std::isfinite(c)) { lo[a] = std::min(lo[a], c); hi[a] = std::max(hi[a], c); }
                }
            auto spread = [](uint32_t v) {  // 10 bits -> every third bit
                v &= 1023u; v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu;
                v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u; return v;
            };
            std::vector<uint32_t> code(n);
            for (uint32_t i = 0; i < n; ++i) {
                uint32_t m = 0;
                for (int a = 0; a < 3; ++a) {
                    const float ext = hi[a] - lo[a];
                    float t = (ext > 0.f && std::isfinite(cen[3 * (size_t)i + a])) ? (cen[3 * (size_t)i + a] - lo[a]) / ext : 0.f;
                    t = std::min(std::max(t, 0.f), 1.f);
                    m |= spread((uint32_t)(t * 1023.f)) << a;
                }
                code[i] = m;
            }
            for (int c = 0; c < 3; ++c)
                std::stable_sort(perm.begin() + start[c], perm.begin() + start[c] + cnt[c], [&](uint32_t x, uint32_t y) { return code[x] < code[y]; });
        }
    }
    PinnedStage<float4>& sph = g_stage_sph;
    sph.resize((size_t)2 * std::max(sc->n_spheres, 1u));
    PinnedStage<rt_material>& mats = g_stage_mat;
    mats.resize(sc->n_materials);
    if (!perm.data() || !sph.data() || !mats.data()) return fail(RT_ERR_CUDA, "out of host memory for the scene staging buffers");
    memcpy(mats.data(), sc->materials, sizeof(rt_material) * sc->n_materials);
    for (uint32_t i = 0; i < sc->n_spheres; ++i) {
        const rt_sphere& s = sc->spheres[i];
        sph[2 * i] = make_float4(s.center[0], s.center[1], s.center[2], s.radius);
        float4 m = make_float4(0, 0, 0, 0);
        memcpy(&m.x, &s.material, 4);
        sph[2 * i + 1] = m;
        for (int a = 0; a < 3; ++a) extent = std::max(extent, std::fabs(s.center[a]) + std::fabs(s.radius));
    }
    g.any_transparent = false;
    g.max_ni = 1.f;
    for (uint32_t i = 0; i < sc->n_materials; ++i) {
        if ((sc->materials[i].flags & RT_HAS_TR) && sc->materials[i].Tr < 1.0f) g.any_transparent = true;
        const float ni = std::fabs(sc->materials[i].Ni);
        const float worst = (ni > 0.f && std::isfinite(ni)) ? std::max(ni, 1.0f / ni) : INFINITY;   // Ni = 0 / NaN: unbounded
        if (sc->materials[i].Tr < 1.0f || !(sc->materials[i].Tr == sc->materials[i].Tr)) g.max_ni = std::max(g.max_ni, worst);
    }
    g.scene_extent = extent;

    // Stream-ordered from here on: the copies queue behind whatever frame is still in flight, nothing waits on the host
    // (a reallocation in ensure() synchronises by itself).  The pinned staging buffers stay alive; ev_stage marks the
    // point where the copies have left them (waited on at the top of the next upload).
    for (RtDevice& d : g.devs) {
        CU(cudaSetDevice(d.device));
        d.ntri = (int)n;
        d.ntiles = cls_tiles[0] + cls_tiles[1] + cls_tiles[2];
        d.cls1 = cls_tiles[0];
        d.cls2 = cls_tiles[0] + cls_tiles[1];
        d.nmat = (int)sc->n_materials;
        d.nspheres = (int)sc->n_spheres;
        d.M_built = 0.f; d.dir_built = 0.f;
        // grow-only device buffers: re-uploading a scene of the same size allocates nothing
        rc = ensure(d.rec, d.cap_rec, (size_t)(d.ntiles + kPadTiles) * kTile * kRecVec); if (rc) return rc;
        rc = ensure(d.tile_box, d.cap_box, (size_t)(d.ntiles + kPadTiles) * 2); if (rc) return rc;
        rc = ensure(d.perm, d.cap_perm, perm.size()); if (rc) return rc;
        CU(cudaMemcpyAsync(d.perm, perm.data(), sizeof(uint32_t) * perm.size(), cudaMemcpyHostToDevice, d.stream));
        rc = ensure(d.triv, d.cap_triv, triv.size()); if (rc) return rc;
        rc = ensure(d.normal_mat, d.cap_nm, nm.size()); if (rc) return rc;
        rc = ensure(d.materials, d.cap_mat, (size_t)4 * sc->n_materials); if (rc) return rc;
        rc = ensure(d.spheres, d.cap_sph, sph.size()); if (rc) return rc;
        CU(cudaMemcpyAsync(d.triv, triv.data(), sizeof(float4) * triv.size(), cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.normal_mat, nm.data(), sizeof(float4) * nm.size(), cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.materials, mats.data(), sizeof(rt_material) * sc->n_materials, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.spheres, sph.data(), sizeof(float4) * sph.size(), cudaMemcpyHostToDevice, d.stream));
        rc = ensure(d.tri_group, d.cap_group, grp.size()); if (rc) return rc;
        CU(cudaMemcpyAsync(d.tri_group, grp.data(), grp.size(), cudaMemcpyHostToDevice, d.stream));
        CU(cudaEventRecord(d.ev_stage, d.stream));
        if (!g_stage_triv.pinned || !g_stage_nm.pinned || !g_stage_perm.pinned || !g_stage_mat.pinned || !g_stage_sph.pinned || !g_stage_group.pinned)
            CU(cudaStreamSynchronize(d.stream));   // pageable fallback: keep the old, synchronous behaviour
    }
    g.scene_ready = true;
    g.frame_ready = false;
    return RT_OK;
}

int rt_render_async(const rt_params* params) { return render_enqueue(params); }

int rt_sync(void) {
    int rc = check_ready();
    if (rc) return rc;
    return sync_all();
}

int rt_render(const rt_params* params) {
    int rc = render_enqueue(params);
    if (rc) return rc;
    return sync_all();
}

int rt_download_framebuffer(float* rgb, int32_t* prim_id) {
    int rc = check_ready();
    if (rc) return rc;
    if (!g.frame_ready) return fail(RT_ERR_STATE, "rt_download_framebuffer before rt_render");
    if (!rgb && !prim_id) return fail(RT_ERR_INVALID, "nothing to download");
    RtDevice& d0 = g.devs[0];
    const rt_params& rp = g.last;
    const uint32_t W = rp.width, H = rp.height, G = (uint32_t)g.world;
    CU(cudaSetDevice(d0.device));
    CU(cudaStreamSynchronize(d0.stream));
    if (rgb) {
        const float* src = (G > 1) ? d0.fb_final : d0.fb_local;
        CU(cudaMemcpy(rgb, src, sizeof(float) * 3 * W * H, cudaMemcpyDeviceToHost));
    }
    if (prim_id) {
        if (!rp.want_prim_id) return fail(RT_ERR_STATE, "prim_id requested but the frame was rendered with want_prim_id == 0");
        const size_t row = (size_t)W * rp.pixelfactor_x * rp.pixelfactor_y;
        if (g.single_process) {
            std::vector<int32_t> tmp((size_t)g.rows_per_rank * row);
            for (RtDevice& d : g.devs) {
                CU(cudaSetDevice(d.device));
                CU(cudaStreamSynchronize(d.stream));
                CU(cudaMemcpy(tmp.data(), d.prim, sizeof(int32_t) * tmp.size(), cudaMemcpyDeviceToHost));
                for (uint32_t ly = 0; (size_t)ly * G + d.rank < H; ++ly)
                    memcpy(prim_id + ((size_t)ly * G + d.rank) * row, tmp.data() + (size_t)ly * row, sizeof(int32_t) * row);
            }
        } else {
            // one process per GPU: only this rank's rows are known here; the others are left as -2
            std::vector<int32_t> tmp((size_t)g.rows_per_rank * row);
            CU(cudaMemcpy(tmp.data(), d0.prim, sizeof(int32_t) * tmp.size(), cudaMemcpyDeviceToHost));
            if (G > 1) for (size_t i = 0; i < (size_t)H * row; ++i) prim_id[i] = -2;
            for (uint32_t ly = 0; (size_t)ly * G + d0.rank < H; ++ly)
                memcpy(prim_id + ((size_t)ly * G + d0.rank) * row, tmp.data() + (size_t)ly * row, sizeof(int32_t) * row);
        }
    }
    return RT_OK;
}

int rt_download_framebuffer_u8(uint8_t* rgb8) {
    int rc = check_ready();
    if (rc) return rc;
    if (!g.frame_ready) return fail(RT_ERR_STATE, "rt_download_framebuffer_u8 before rt_render");
    if (!rgb8) return fail(RT_ERR_INVALID, "rgb8 is NULL");
    RtDevice& d0 = g.devs[0];
    const size_t n = (size_t)3 * g.last.width * g.last.height;
    CU(cudaSetDevice(d0.device));
    rc = ensure(d0.fb_u8, d0.cap_u8, n);
    if (rc) return rc;
    const float* src = (g.world > 1) ? d0.fb_final : d0.fb_local;
    k_quantise<<<d0.num_sms * 4, 256, 0, d0.stream>>>(src, d0.fb_u8, n);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rgb8, d0.fb_u8, n, cudaMemcpyDeviceToHost, d0.stream));
    CU(cudaStreamSynchronize(d0.stream));
    return RT_OK;
}

static int rt_trace_impl(const rt_params* rp, int n, const float* origins, const float* dests, float* rgb, int32_t* prim_id, float* hit) {
    int rc = check_ready();
    if (rc) return rc;
    if (!g.scene_ready) return fail(RT_ERR_STATE, "rt_trace before rt_upload_scene");
    rc = validate_params(rp, false);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!origins || !dests || !rgb))) return fail(RT_ERR_INVALID, "bad rt_trace arguments");
    if (n == 0) return RT_OK;
    if ((size_t)n > kMaxChunkSamples) return fail(RT_ERR_INVALID, "rt_trace: at most %u rays per call", kMaxChunkSamples);
    RtDevice& d = g.devs[0];
    CU(cudaSetDevice(d.device));
    (void)cudaGetLastError();   // a stale code from an earlier, already reported failure must not be blamed on this call's launches
    d.kev_kind.clear();
    d.used_graph = false;
    float M = magnitude_bound(*rp, origins, 3 * n);
    rc = build_records(d, M, direction_bound(*rp, false, origins, dests, n)); if (rc) return rc;
    rc = ensure_chunk_state(d, (size_t)n, false, 0); if (rc) return rc;
    rc = ensure_counters(d, 1); if (rc) return rc;
    // grow-only buffers, pinned staging: a call allocates nothing once the sizes have been seen
    const bool want_hit = prim_id || hit;
    rc = ensure(d.trace_in, d.cap_trace_in, (size_t)2 * n); if (rc) return rc;
    rc = ensure(d.hit0, d.cap_hit0, (size_t)n); if (rc) return rc;
    g_stage_rays.resize((size_t)2 * n);
    g_stage_out.resize((size_t)2 * n);
    if (!g_stage_rays.data() || !g_stage_out.data()) return fail(RT_ERR_CUDA, "out of host memory for the rt_trace staging buffers");
    float4* in = g_stage_rays.data();
    for (int i = 0; i < n; ++i) {
        in[2 * i] = make_float4(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2], 0.f);
        in[2 * i + 1] = make_float4(dests[3 * i], dests[3 * i + 1], dests[3 * i + 2], 0.f);
    }
    d.frame_chunks = 1;
    d.pencil_used = 0;
    FrameParams P;
    fill_common(P, d, *rp, eps_r_for(M), d.counters);
    P.trace_api = 1;
    P.nsamples = (uint32_t)n;
    P.nslots = (uint32_t)n;
    P.G = 1;
    float4* out = g_stage_out.data();
    // a handful of rays (the drop-in performRayTracing call is n = 1): everything in one launch of one CTA, exact tests only
    const bool small = g.small_trace && n <= kSmallTraceRays && (double)n * (double)std::max(d.ntri, 1) <= kSmallTraceTests;
    // one H2D copy in, the wavefront (the level-0 hit records are copied aside on the device before the bounces
    // overwrite them), two D2H copies out -- all in stream order
    auto body = [&]() -> int {
        CU(cudaMemcpyAsync(d.trace_in, in, sizeof(float4) * 2 * n, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemsetAsync(d.counters, 0, sizeof(uint32_t) * kCntWords, d.stream));
        k_init_trace<<<(n + 255) / 256, 256, 0, d.stream>>>(d.trace_in, n, d.ray_o, d.ray_d, d.thr, d.acc);
        CU(cudaGetLastError());
        if (small) {
            // single-launch path (rt_kernels.cuh: k_trace_small): one CTA carries the batch through every level
            const bool bounces = (P.features & (RT_REFLECTION | RT_REFRACTION)) != 0;
            const int levels = bounces ? std::min(P.max_lvl + 1, kMaxLevels - 2) : 1;
            LaunchTimer t(d, kKindTrace);
            k_trace_small<<<1, kSmallThreads, 0, d.stream>>>(P, levels, want_hit ? d.hit0 : nullptr, g.any_transparent ? 1 : 0);
            CU(cudaGetLastError());
        } else {
            const int levels = run_wavefront(d, P, want_hit ? d.hit0 : nullptr);
            if (levels < 0) return levels;
        }
        CU(cudaMemcpyAsync(out, d.acc, sizeof(float4) * n, cudaMemcpyDeviceToHost, d.stream));
        if (want_hit) CU(cudaMemcpyAsync(out + n, d.hit0, sizeof(float4) * n, cudaMemcpyDeviceToHost, d.stream));
        return RT_OK;
    };
    // small batches (the drop-in performRayTracing call is n = 1) replay a captured graph: copies and launches in one submission
    const bool use_graph = g_stage_rays.pinned && g_stage_out.pinned &&
                           (g.graph_mode > 0 || (g.graph_mode < 0 && (double)n * (double)std::max(d.ntri, 1) <= kGraphMaxTests && !getenv("RT_B200_LAUNCHLOG")));
    if (!use_graph) {
        rc = body(); if (rc) return rc;
    } else {
        std::vector<uint8_t> key;
        KeyWriter kw{key};
        rt_params rk = *rp;
        memset(rk.corners, 0, sizeof(rk.corners)); rk.width = rk.height = rk.pixelfactor_x = rk.pixelfactor_y = 0; rk.want_prim_id = 0;   // ignored by rt_trace
        kw.put(rk); kw.put(d.rec_gen); kw.put(n); kw.put(want_hit); kw.put(P.eps_r); kw.put(small);
        const void* ptrs[] = {d.rec, d.triv, d.normal_mat, d.materials, d.spheres, d.tile_box, d.super_box, d.always_list, d.ray_o, d.ray_d, d.thr, d.acc,
                              d.hit, d.lit, d.q_ray, d.q_hit, d.key, d.counters, d.trace_in, d.hit0, in, out};
        kw.put(ptrs);
        const int cfg[] = {g.scan.rp, g.scan.j, g.scan.minb, (int)g.tile_culling, d.num_sms, (int)g.any_transparent};
        kw.put(cfg);
        if (!d.trace_graph || key != d.trace_key) {
            if (d.trace_graph) { cudaGraphExecDestroy(d.trace_graph); d.trace_graph = nullptr; }
            d.trace_key.clear();
            CU(cudaStreamBeginCapture(d.stream, cudaStreamCaptureModeThreadLocal));
            d.capturing = true;
            rc = body();
            d.capturing = false;
            cudaGraph_t graph = nullptr;
            const cudaError_t e = cudaStreamEndCapture(d.stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(RT_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
            const cudaError_t e2 = cudaGraphInstantiate(&d.trace_graph, graph, 0);
            cudaGraphDestroy(graph);
            if (e2 != cudaSuccess) { d.trace_graph = nullptr; return fail(RT_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e2)); }
            d.trace_key = key;
            d.launches_in_trace_graph = (uint32_t)d.kev_kind.size();
        }
        d.kev_kind.assign(d.launches_in_trace_graph, -1);
        CU(cudaGraphLaunch(d.trace_graph, d.stream));
        d.used_graph = true;
    }
    CU(cudaStreamSynchronize(d.stream));
    for (int i = 0; i < n; ++i) {
        rgb[3 * i] = out[i].x; rgb[3 * i + 1] = out[i].y; rgb[3 * i + 2] = out[i].z;
        if (prim_id) memcpy(&prim_id[i], &out[n + i].w, 4);
        if (hit) { hit[3 * i] = out[n + i].x; hit[3 * i + 1] = out[n + i].y; hit[3 * i + 2] = out[n + i].z; }
    }
    return RT_OK;
}

int rt_trace(const rt_params* rp, int n, const float* origins, const float* dests, float* rgb, int32_t* prim_id, float* hit) {
    // the framebuffer of the last frame is not touched by rt_trace (it stays downloadable); the counters are
    g.stats_ready = false;
    const int rc = rt_trace_impl(rp, n, origins, dests, rgb, prim_id, hit);
    if (rc != RT_OK)
        for (RtDevice& d : g.devs) { d.key_dirty = true; d.capturing = false; }
    else { g.stats_ready = true; g.stats_of_trace = true; g.last_trace = *rp; g.last_trace_n = n; }
    return rc;
}

int rt_set_option(int option, int value) {
    if (option == RT_OPT_TILE_CULLING) { g.tile_culling = value != 0; return RT_OK; }
    if (option == RT_OPT_PENCIL) { g.pencil = value != 0; return RT_OK; }
    if (option == RT_OPT_PENCIL_ANY) { g.pencil_any = value != 0; return RT_OK; }
    if (option == RT_OPT_PENCIL_REFLECT) { g.pencil_reflect = value != 0; return RT_OK; }
    if (option == RT_OPT_PENCIL_REFLECT) { g.pencil_reflect = value != 0; return RT_OK; }
    if (option == RT_OPT_SMALL_TRACE) { g.small_trace = value != 0; return RT_OK; }
    if (option == RT_OPT_PENCIL_THREAD) { g.pencil_thread = value < 0 ? 0 : (value > 2 ? 2 : value); return RT_OK; }
    if (option == RT_OPT_GRAPH) { g.graph_mode = value < 0 ? -1 : (value != 0); return RT_OK; }
    return fail(RT_ERR_INVALID, "unknown option %d", option);
}

int rt_probe_fp32_peak(float* tflops) {
    int rc = check_ready();
    if (rc) return rc;
    if (!tflops) return fail(RT_ERR_INVALID, "tflops is NULL");
    RtDevice& d = g.devs[0];
    CU(cudaSetDevice(d.device));
    const int grid = d.num_sms * 8, iters = 8192;
    float2* buf = nullptr;
    CU(cudaMalloc(&buf, sizeof(float2) * (size_t)grid * 256));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    k_fp32_peak_probe<true><<<grid, 256, 0, d.stream>>>(buf, 1.0f, iters / 8);   // warm-up (clock ramp)
    float best = 0.f;
    for (int rep = 0; rep < 6; ++rep) {   // packed and scalar FMAs, best of three each
        CU(cudaEventRecord(e0, d.stream));
        if (rep & 1) k_fp32_peak_probe<true><<<grid, 256, 0, d.stream>>>(buf, 1.0f, iters);
        else k_fp32_peak_probe<false><<<grid, 256, 0, d.stream>>>(buf, 1.0f, iters);
        CU(cudaEventRecord(e1, d.stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = (double)grid * 256 * (double)iters * 16 * 4;   // 16 FFMA2 per iteration, 4 flop each
        best = std::max(best, (float)(flop / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    *tflops = best;
    return RT_OK;
}

int rt_get_stats(rt_stats* out) {
    int rc = check_ready();
    if (rc) return rc;
    if (!out) return fail(RT_ERR_INVALID, "out is NULL");
    if (!g.stats_ready) return fail(RT_ERR_STATE, "rt_get_stats before a completed rt_render / rt_trace");
    rc = sync_all();
    if (rc) return rc;
    rc = collect_stats();
    if (rc) return rc;
    *out = g.stats;
    return RT_OK;
}

int rt_event_record(int slot) {
    int rc = check_ready();
    if (rc) return rc;
    if (slot < 0 || slot >= 16) return fail(RT_ERR_INVALID, "event slot out of range");
    for (RtDevice& d : g.devs) {
        CU(cudaSetDevice(d.device));
        CU(cudaEventRecord(d.ev[slot], d.stream));
    }
    return RT_OK;
}

int rt_event_elapsed_ms(int a, int b, float* ms) {
    int rc = check_ready();
    if (rc) return rc;
    if (a < 0 || a >= 16 || b < 0 || b >= 16 || !ms) return fail(RT_ERR_INVALID, "bad event arguments");
    float worst = 0.f;
    for (RtDevice& d : g.devs) {
        CU(cudaSetDevice(d.device));
        CU(cudaEventSynchronize(d.ev[b]));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, d.ev[a], d.ev[b]));
        worst = std::max(worst, t);
    }
    *ms = worst;
    return RT_OK;
}

}  // extern "C"
