// map2d.cu — host side of libmap2d_b200.so: the Map2D object behind the C-ABI of include/map2d_b200.h.
//
// Mirrors the reference classes Map2DCPU / MultiBandMap2DCPU (Map2DFusion/Map2DCPU.cpp, MultiBandMap2DCPU.cpp):
// prepare() lays out the tile grid; feed()/feed_batch() decide every frame's tile rectangle on the host in FP64
// (geom.h, the same code the bounds kernel runs), allocate first-touch tiles from a slab pool in HBM, build the
// group's tile-centric work list and enqueue the fusion kernels on the handle's stream.  Frames are fused in
// groups of up to K (m2d_config.batch_frames); a single feed() is a group of one.  There is no CPU fallback:
// without a CUDA device every call fails loudly.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/map2d_b200.h"
#include "geom.h"
#include "kernels.cuh"
#include "bounds.h"

int m2d_write_png(const char* path, const uint8_t* bgr_or_bgra, int w, int h, int channels);  // png.cpp

using namespace m2d;

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char buf_[512];                                                                           \
            snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            err = buf_;                                                                               \
            return M2D_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)
#define LAUNCHKS(kind, strm, call)                            \
    do {                                                      \
        cudaEvent_t pe0_ = nullptr, pe1_ = nullptr;           \
        if (profiling) {                                      \
            CU(cudaEventCreate(&pe0_));                       \
            CU(cudaEventCreate(&pe1_));                       \
            CU(cudaEventRecord(pe0_, strm));                  \
        }                                                     \
        CU(call);                                             \
        launches++;                                           \
        if (profiling) {                                      \
            CU(cudaEventRecord(pe1_, strm));                  \
            prof.push_back(ProfRec{kind, pe0_, pe1_});        \
        }                                                     \
    } while (0)
#define LAUNCHK(kind, call) LAUNCHKS(kind, stream, call)
#define LAUNCH(call) LAUNCHK(M2D_K_MISC, call)

struct ProfRec { int kind; cudaEvent_t e0, e1; };

// While an ingest worker exists it shares the handle with the caller's threads: every entry point then serialises on
// the handle's mutex (no cost otherwise; a handle without ingest stays single-threaded by contract).
#define API_LOCK(h)                                         \
    std::unique_lock<std::recursive_mutex> api_lk_;         \
    if ((h) && (h)->ingest) api_lk_ = std::unique_lock<std::recursive_mutex>((h)->api)

// Device + pinned buffers of one in-flight group.  Two contexts alternate so that the host can prepare group g+1
// (bounds, tile allocation, work lists, H2D copies) while the GPU fuses group g.
struct GroupCtx {
    cudaEvent_t copied = nullptr;     // host frames of the group have landed in d_raw (copy stream)
    cudaEvent_t staged = nullptr;     // order-independent stages (warp / pyramid) finished on `stage`
    cudaEvent_t done = nullptr;       // recorded after the group's last kernel
    cudaEvent_t decided = nullptr;    // weights-first multi-band: winners + need flags of the group are known
    cudaStream_t stage = nullptr;     // per-context stream: stages of group g+1 overlap the select of group g
    bool busy = false;
    int frames = 0;
    uint8_t* h_blob = nullptr;        // pinned: [FrameJob x K | TileWork x T | TileEntry x E]
    uint8_t* d_blob = nullptr;
    size_t blob_cap = 0;
    uint8_t* d_raw = nullptr;         // K x sw*sh*3 staging for host frames
    size_t raw_cap = 0;
    uint8_t* d_scratch = nullptr;     // multi-band pyramids of the group
    size_t scratch_cap = 0;
};

// Ingest seam: bounded drop-oldest frame queue in front of feed(), pinned slots, one worker thread.
struct IngestItem { int slot; double pose[7]; };
struct Ingest {
    std::mutex mu;
    std::condition_variable cv, cv_idle;
    std::deque<IngestItem> q;          // frames waiting, oldest first
    std::vector<int> free_slots;
    uint8_t* ring = nullptr;           // pinned: n_slots x slot_bytes of packed BGR8
    size_t slot_bytes = 0;
    int capacity = 0, n_slots = 0, w = 0, h = 0;
    bool paused = false, stop = false, busy = false;
    uint64_t pushed = 0, dropped = 0, fed = 0, fused = 0;
    int last_rc = M2D_OK;
    std::thread worker;
};
constexpr int kIngestBatch = 16;       // frames the worker hands to feed() at once
constexpr int kIngestSpare = 4;        // slots for producers that are mid-copy

struct m2d_map {
    // Multi-device handle (m2d_create_multi): ONE host process, one sub-handle per GPU, tiles sharded by spatial ownership
    // (block-cyclic strips, m2d_config.shard_axis / shard_span).  The top handle owns no CUDA resources; prepare / feed /
    // sync / reset / save / getters are routed to the sub-handles.  Peer access is enabled between the devices, so a
    // device-resident frame is sampled in place over NVLink by every GPU that owns tiles under it, and the save gathers raw
    // tiles to the first device with plain peer loads.
    std::vector<m2d_map*> subs;
    void mirror() { if (!subs.empty()) { g = subs[0]->g; valid = subs[0]->valid; min_z = subs[0]->min_z; max_z = subs[0]->max_z; length_pixel = subs[0]->length_pixel;
                                          org_x = subs[0]->org_x; org_y = subs[0]->org_y; stats = subs[0]->stats; memcpy(last_rect, subs[0]->last_rect, sizeof last_rect); } }
    int type = 0;
    m2d_config cfg{};
    int band_num = 5, levels = 6;
    bool valid = false;
    std::string err;
    uint64_t launches = 0;
    bool profiling = false;
    std::vector<ProfRec> prof;

    // Map2DPrepare + Map2DCPUData
    GridGeom g{};
    double min_z = 0, max_z = 0, length_pixel = 0;
    int org_x = 0, org_y = 0;       // absolute tile coordinate of grid slot (0,0); moves under spreadMap
    int shard_origin = 0;           // absolute tile coordinate at which strip 0 starts (m2d_set_shard)
    std::vector<uint8_t*> table;    // tile state pointer per grid slot (w*h), NULL = untouched / not owned
    // host work-list scratch, reused across groups (no per-group allocation or hashing on the feed path)
    std::vector<uint8_t> changed;           // per grid slot: the reference's Ele::Ischanged (set by feed, cleared by poll)
    std::vector<int> slot_work;             // per grid slot: index into the current group's TileWork list ...
    std::vector<uint32_t> slot_epoch;       // ... valid only when the epoch matches
    uint32_t group_epoch = 0;
    std::vector<std::vector<TileEntry>> per_tile_pool;
    int last_rect[4] = {-1, -1, -1, -1};

    // tile pool
    TileLayout lay{};
    size_t tile_bytes = 0;
    std::vector<void*> chunks;
    std::vector<uint8_t*> free_tiles;
    size_t tiles_in_use = 0;
    // pool filler: a helper thread keeps at least `pool_low` tiles free by allocating the next slab AHEAD of need, so
    // that a streaming feed() never waits for a cudaMalloc (5-7 ms for a 128 MiB slab: it was the whole p99 of cfg5)
    std::mutex pool_mu;
    std::condition_variable pool_cv;
    std::thread pool_thread;
    bool pool_stop = false, pool_failed = false;
    size_t pool_low = 0;

    // per-size weight images
    int wimg_w = 0, wimg_h = 0;
    uint8_t* d_alpha = nullptr;
    float* d_wimg = nullptr;

    uint8_t* d_collapse = nullptr;  // cached buffers of get_image()/save()
    size_t collapse_cap = 0;

    static constexpr int kMaxCtx = 8;
    static constexpr int kMaxGroup = 1024;   // frames per group (the winner map holds 16-bit entry indices per tile)
    int kCtx = 4;                   // group contexts in flight (M2D_CTX env overrides, for tuning)
    bool weights_first = true;      // multi-band decides winners from the weight pyramids first, then warps and filters the
                                    // image only where a winner needs it (kernels.cu "WEIGHTS-FIRST variant"); M2D_SPARSE=0
                                    // selects the dense pipeline (every frame fully warped and filtered) for A/B runs
    cudaStream_t decide_stream = nullptr;  // chain of the groups' bounds + decide stages (tile weights, cmin), ahead of the Laplacian chain
    cudaEvent_t dense_done = nullptr;      // last dense-pipeline select on the handle's stream (small groups), see run_group
    bool dense_pending = false;
    bool weight_cull = true;        // bound-based culling of (frame, cell) pairs before any weight is computed (bounds.h);
                                    // M2D_WCULL=0 treats every covering frame as competitive (A/B runs, parity sweeps)
    cudaEvent_t input_event = nullptr;   // m2d_set_input_event: the frames of the next feed call are valid once it fires
    bool use_tma = true;            // M2D_TMA=0: pyrDown 0 -> 1 with the register-window kernel instead of TMA-staged patches (A/B)
    double scratch_budget = 12e9;   // bytes of group scratch per context that the default group size aims at (M2D_SCRATCH_GB)
    int sm_count = 148;
    GroupCtx ctx[kMaxCtx];
    int ctx_next = 0;

    Ingest* ingest = nullptr;
    std::recursive_mutex api;       // serialises the ingest worker with API calls (taken only while ingest is open)

    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;  // H2D staging of host frames overlaps the previous group's kernels

    unsigned long long* d_stats = nullptr;  // [0..8] level wins, [16] footprint, [17] weighted wins
    m2d_stats stats{};

    // Map2DRender (type 4): one batch -> one blended canvas.  Grow-only device buffers, the last result stays in HBM.
    uint8_t* rnd_wimg = nullptr;            // 8-bit weight image of the camera size (Map2DRender.cpp:507-529)
    int rnd_wimg_w = 0, rnd_wimg_h = 0;
    uint8_t* rnd_canvas = nullptr;          // canvas pyramids (planar int16 Laplacians + weights) and the outputs
    size_t rnd_canvas_cap = 0;
    uint8_t* rnd_scratch = nullptr;         // sub-image pyramids of the frames of one chunk
    size_t rnd_scratch_cap = 0;
    uint8_t* rnd_raw = nullptr;             // staging of host frames (one chunk)
    size_t rnd_raw_cap = 0;
    uint8_t* rnd_jobs = nullptr;            // RenderJob array of one chunk
    size_t rnd_jobs_cap = 0;
    bool rnd_have = false;
    int rnd_w = 0, rnd_h = 0, rnd_bands = 0, rnd_tx0 = 0, rnd_ty0 = 0;   // tx0, ty0: ABSOLUTE tile coordinate of the canvas origin
    size_t rnd_off16 = 0, rnd_off8 = 0, rnd_offmask = 0;
    int render_frames(int n, const uint8_t* base, size_t frame_stride, int w, int h, size_t stride, const double* poses, bool on_device,
                      int* result);

    int init();
    void release();
    int prepare(const double* plane, const double* cam, int n, const double* poses);
    int spread(double xmin, double ymin, double xmax, double ymax);
    // frame i = ptrs ? ptrs[i] : base + i * frame_stride
    int feed_frames(int n, const uint8_t* base, size_t frame_stride, int w, int h, size_t stride, const double* poses,
                    bool on_device, int* result, const uint8_t* const* ptrs = nullptr);
    int run_group(int n, const uint8_t* base, size_t frame_stride, int w, int h, size_t stride, const double* poses,
                  bool on_device, int* result, const uint8_t* const* ptrs, const uint8_t* pull_base = nullptr);
    // Pinned host frames (m2d_alloc_host / cudaHostAlloc: device-visible under UVA) are not staged whole: weighted mode samples
    // them in place over PCIe, the weights-first multi-band pipeline pulls only the chunks its winners need (kernels_wf.cu
    // mbs_mark / mbs_pull).  M2D_ZEROCOPY=0 keeps the cudaMemcpyAsync staging of every frame (A/B runs).
    bool zero_copy = true;
    bool pull_poison = false;       // M2D_PULL_POISON=1 (tests): fill the staging slots with 0xA5 before every pull
    int ensure_weight_images(int w, int h);
    int alloc_tile(uint8_t** out);
    int reserve_tiles(size_t n);
    void pool_filler();
    int grow(void** p, size_t* cap, size_t need, bool pinned);
    int group_size(int w, int h, bool on_device) const;
    bool owns(int tx, int ty) const;
    bool tile_bbox(int& x0, int& y0, int& x1, int& y1) const;
    int get_image(uint8_t* out, int* w, int* h, int* channels, int* tmx, int* tmy);
    int collapse_window(uint8_t* out, bool out_on_device, const int win[4], const int crop[4], int* w, int* h, int* channels);
    int multi_get_image(uint8_t* out, int* w, int* h, int* channels, int* tmx, int* tmy);
    int queue_size();
    int sync();
    int reset();
};

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

bool m2d_map::owns(int tx, int ty) const {
    if (cfg.shard_count <= 1) return true;
    int a = ((cfg.shard_axis == 0) ? tx + org_x : ty + org_y) - shard_origin;
    int span = cfg.shard_span > 0 ? cfg.shard_span : 1;
    int s = floordiv(a, span) % cfg.shard_count;
    if (s < 0) s += cfg.shard_count;
    return s == cfg.shard_rank;
}

int m2d_map::init() {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)";
        return M2D_ERR_CUDA;
    }
    CU(cudaSetDevice(cfg.device));
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;
    CU(cudaMalloc(&d_stats, 32 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(d_stats, 0, 32 * sizeof(unsigned long long), stream));
    if (const char* e = getenv("M2D_CTX")) kCtx = std::max(2, std::min(atoi(e), (int)kMaxCtx));
    if (const char* e = getenv("M2D_SPARSE")) weights_first = atoi(e) != 0;
    if (const char* e = getenv("M2D_WCULL")) weight_cull = atoi(e) != 0;
    if (const char* e = getenv("M2D_TMA")) use_tma = atoi(e) != 0;
    if (const char* e = getenv("M2D_ZEROCOPY")) zero_copy = atoi(e) != 0;
    if (const char* e = getenv("M2D_PULL_POISON")) pull_poison = atoi(e) != 0;
    if (const char* e = getenv("M2D_SCRATCH_GB")) scratch_budget = std::max(0.25, atof(e)) * 1e9;
    CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, cfg.device));
    CU(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&decide_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&dense_done, cudaEventDisableTiming));
    for (int i = 0; i < kCtx; i++) {
        CU(cudaEventCreateWithFlags(&ctx[i].done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx[i].copied, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx[i].staged, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx[i].decided, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&ctx[i].stage, cudaStreamNonBlocking));
    }
    levels = (type == M2D_TYPE_MULTIBAND) ? band_num + 1 : 1;
    if (type == M2D_TYPE_MULTIBAND) {
        lay = make_tile_layout(levels);
        tile_bytes = lay.bytes;
    } else {
        tile_bytes = (size_t)kEle * kEle * 4;   // (TypeRender keeps no tiles: its state is the canvas of the last batch)
    }
    return M2D_OK;
}

void m2d_map::release() {
    if (!subs.empty()) {
        for (m2d_map* sub : subs) { sub->release(); delete sub; }
        subs.clear();
        return;
    }
    cudaSetDevice(cfg.device);
    if (pool_thread.joinable()) {
        { std::lock_guard<std::mutex> lk(pool_mu); pool_stop = true; }
        pool_cv.notify_all();
        pool_thread.join();
    }
    if (stream) cudaStreamSynchronize(stream);
    for (void* c : chunks) cudaFree(c);
    chunks.clear();
    free_tiles.clear();
    if (d_alpha) cudaFree(d_alpha);
    if (d_wimg) cudaFree(d_wimg);
    for (int i = 0; i < kMaxCtx; i++) {
        GroupCtx& c = ctx[i];
        if (c.done) cudaEventDestroy(c.done);
        if (c.copied) cudaEventDestroy(c.copied);
        if (c.staged) cudaEventDestroy(c.staged);
        if (c.decided) cudaEventDestroy(c.decided);
        if (c.stage) { cudaStreamSynchronize(c.stage); cudaStreamDestroy(c.stage); }
        if (c.h_blob) cudaFreeHost(c.h_blob);
        if (c.d_blob) cudaFree(c.d_blob);
        if (c.d_raw) cudaFree(c.d_raw);
        if (c.d_scratch) cudaFree(c.d_scratch);
    }
    if (d_stats) cudaFree(d_stats);
    if (d_collapse) cudaFree(d_collapse);
    if (rnd_wimg) cudaFree(rnd_wimg);
    if (rnd_canvas) cudaFree(rnd_canvas);
    if (rnd_scratch) cudaFree(rnd_scratch);
    if (rnd_raw) cudaFree(rnd_raw);
    if (rnd_jobs) cudaFree(rnd_jobs);
    for (ProfRec& r : prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
    if (decide_stream) { cudaStreamSynchronize(decide_stream); cudaStreamDestroy(decide_stream); }
    if (dense_done) cudaEventDestroy(dense_done);
    if (own_stream && stream) cudaStreamDestroy(stream);
}

// Map2DPrepare::prepare (Map2D.cpp:32-49) + Map2DCPUData::prepare (Map2DCPU.cpp:44-92 / MultiBandMap2DCPU.cpp:199-255)
int m2d_map::prepare(const double* plane7, const double* cam, int n, const double* poses) {
    if (!subs.empty()) {
        int rc = M2D_OK;
        for (m2d_map* sub : subs) { int r = sub->prepare(plane7, cam, n, poses); if (r != M2D_OK) { rc = r; err = sub->err; } }
        mirror();
        return rc;
    }
    if (n <= 0 || !poses || cam[0] <= 0 || cam[1] <= 0 || cam[2] == 0 || cam[3] == 0) return M2D_REJECTED;
    // cv::remap keeps source coordinates in 16-bit maps (saturate_cast<short>); the kernels rely on sw, sh <= 32767
    if (cam[0] > 32767 || cam[1] > 32767) { err = "camera larger than 32767 px: unsupported (OpenCV's remap maps are 16-bit)"; return M2D_ERR_UNSUPPORTED; }
    GridGeom ng{};
    ng.cam_w = cam[0]; ng.cam_h = cam[1]; ng.cx = cam[4]; ng.cy = cam[5];
    ng.fxinv = 1. / cam[2]; ng.fyinv = 1. / cam[3];
    ng.plane_inv = pose_inverse(pose_from7(plane7));
    Vec3 mx{-1e10, -1e10, -1e10}, mn{1e10, 1e10, 1e10};
    for (int i = 0; i < n; i++) {
        Pose p = pose_mul(ng.plane_inv, pose_from7(poses + 7 * i));
        const Vec3& t = p.t;
        mx.x = t.x > mx.x ? t.x : mx.x; mx.y = t.y > mx.y ? t.y : mx.y; mx.z = t.z > mx.z ? t.z : mx.z;
        mn.x = t.x < mn.x ? t.x : mn.x; mn.y = t.y < mn.y ? t.y : mn.y; mn.z = t.z < mn.z ? t.z : mn.z;
    }
    if (mn.z * mx.z <= 0) return M2D_REJECTED;
    double hgt;
    if (type == M2D_TYPE_MULTIBAND) hgt = (mx.z > 0) ? mx.z : -mn.z;
    else hgt = (mn.z > 0) ? mn.z : -mx.z;
    double lx = (ng.cam_w - ng.cx) * ng.fxinv - (0 - ng.cx) * ng.fxinv;
    double ly = (ng.cam_h - ng.cy) * ng.fyinv - (0 - ng.cy) * ng.fyinv;
    double radius = 0.5 * hgt * sqrt((lx * lx + ly * ly));
    double lp = 0;
    if (type == M2D_TYPE_MULTIBAND) lp = cfg.resolution;
    if (!lp) {
        lp = 2 * radius / sqrt(ng.cam_w * ng.cam_w + ng.cam_h * ng.cam_h);
        lp /= cfg.scale;
    }
    mn.x = mn.x - radius; mn.y = mn.y - radius;
    mx.x = mx.x + radius; mx.y = mx.y + radius;
    Vec3 c{0.5 * (mn.x + mx.x), 0.5 * (mn.y + mx.y), 0.5 * (mn.z + mx.z)};
    mn = Vec3{2 * mn.x - c.x, 2 * mn.y - c.y, 2 * mn.z - c.z};
    mx = Vec3{2 * mx.x - c.x, 2 * mx.y - c.y, 2 * mx.z - c.z};
    double ele = kEle * lp;
    int w = (int)ceil((mx.x - mn.x) / ele), h = (int)ceil((mx.y - mn.y) / ele);
    if (w <= 0 || h <= 0 || (long long)w * h > (1ll << 26)) return M2D_REJECTED;
    mx.x = mn.x + ele * w;
    mx.y = mn.y + ele * h;
    ng.min_x = mn.x; ng.min_y = mn.y; ng.max_x = mx.x; ng.max_y = mx.y;
    ng.ele_size = ele; ng.ele_size_inv = 1. / ele; ng.length_pixel_inv = 1. / lp;
    ng.w = w; ng.h = h;

    CU(cudaSetDevice(cfg.device));
    if (valid) {  // a second prepare() starts a fresh map, like the reference swapping in new Prepare/Data objects
        int rc = reset();
        if (rc != M2D_OK) return rc;
    }
    g = ng;
    min_z = mn.z; max_z = mx.z; length_pixel = lp;
    org_x = org_y = 0;
    table.assign((size_t)w * h, nullptr);
    slot_work.assign((size_t)w * h, -1);
    slot_epoch.assign((size_t)w * h, 0);
    changed.assign((size_t)w * h, 0);
    last_rect[0] = last_rect[1] = last_rect[2] = last_rect[3] = -1;
    valid = true;
    rnd_have = false;
    if (type == M2D_TYPE_RENDER) return M2D_OK;   // no tiles: Map2DRenderData::_data is never filled (Map2DRender.cpp:479-760)
    // Pre-reserve pool slabs for a quarter of the prepared grid (prepare() doubles the pose bbox about its centre,
    // so ~1/4 of the grid is what the prepare-frames actually cover), capped at 2 GiB; failure here is not fatal.
    size_t want = std::min<size_t>(((size_t)w * h + 3) / 4, ((size_t)2 << 30) / tile_bytes);
    if (reserve_tiles(want) != M2D_OK) err.clear();
    if (!pool_thread.joinable()) {
        // keep the tiles of about two frames free at all times (a frame region is (W/256 + 2) x (H/256 + 2) tiles at scale 1)
        pool_low = std::max<size_t>(32, 2 * (size_t)(cam[0] / kEle + 2) * (size_t)(cam[1] / kEle + 2));
        pool_thread = std::thread([this] { pool_filler(); });
    }
    return M2D_OK;
}

// spreadMap — Map2DCPU.cpp:339-382: grow the grid (never shrinks) and re-index the tile table; no pixel moves and,
// because kernels address tiles by pointer, nothing on the device changes.
// The geometry half of spreadMap (no tile table): new extent in whole tiles around the old origin.
static int grow_geom(GridGeom& g, int& org_x, int& org_y, double xmin, double ymin, double xmax, double ymax,
                     int* dx, int* dy) {
    int xminInt = (int)floor((xmin - g.min_x) * g.ele_size_inv), yminInt = (int)floor((ymin - g.min_y) * g.ele_size_inv);
    int xmaxInt = (int)ceil((xmax - g.min_x) * g.ele_size_inv), ymaxInt = (int)ceil((ymax - g.min_y) * g.ele_size_inv);
    xminInt = std::min(xminInt, 0); yminInt = std::min(yminInt, 0);
    xmaxInt = std::max(xmaxInt, g.w); ymaxInt = std::max(ymaxInt, g.h);
    int nw = xmaxInt - xminInt, nh = ymaxInt - yminInt;
    if (nw <= 0 || nh <= 0 || (long long)nw * nh > (1ll << 26)) return M2D_REJECTED;
    double nminx = g.min_x + g.ele_size * xminInt, nminy = g.min_y + g.ele_size * yminInt;
    double nmaxx = nminx + nw * g.ele_size, nmaxy = nminy + nh * g.ele_size;
    g.min_x = nminx; g.min_y = nminy; g.max_x = nmaxx; g.max_y = nmaxy;
    g.w = nw; g.h = nh;
    org_x += xminInt; org_y += yminInt;
    *dx = xminInt; *dy = yminInt;
    return M2D_OK;
}

int m2d_map::spread(double xmin, double ymin, double xmax, double ymax) {
    const int ow = g.w, oh = g.h;
    int xminInt = 0, yminInt = 0;
    { int rc = grow_geom(g, org_x, org_y, xmin, ymin, xmax, ymax, &xminInt, &yminInt); if (rc != M2D_OK) return rc; }
    const int nw = g.w, nh = g.h;
    std::vector<uint8_t*> nt((size_t)nw * nh, nullptr);
    std::vector<int> nsw((size_t)nw * nh, -1);
    std::vector<uint32_t> nse((size_t)nw * nh, 0);
    std::vector<uint8_t> nch((size_t)nw * nh, 0);
    for (int x = 0; x < ow; x++)
        for (int y = 0; y < oh; y++) {
            size_t o = (size_t)y * ow + x, n = (size_t)(x - xminInt) + (size_t)(y - yminInt) * nw;
            nt[n] = table[o]; nsw[n] = slot_work[o]; nse[n] = slot_epoch[o]; nch[n] = changed[o];
        }
    table.swap(nt); slot_work.swap(nsw); slot_epoch.swap(nse); changed.swap(nch);
    return M2D_OK;
}

int m2d_map::ensure_weight_images(int w, int h) {
    if (wimg_w == w && wimg_h == h) return M2D_OK;
    CU(cudaStreamSynchronize(stream));
    if (d_alpha) { CU(cudaFree(d_alpha)); d_alpha = nullptr; }
    if (d_wimg) { CU(cudaFree(d_wimg)); d_wimg = nullptr; }
    if (type == M2D_TYPE_MULTIBAND) CU(cudaMalloc(&d_wimg, (size_t)w * h * sizeof(float)));
    else CU(cudaMalloc(&d_alpha, (size_t)w * h + 16));
    LAUNCH(launch_weight_images(w, h, cfg.weight_type, d_alpha, d_wimg, stream));
    CU(cudaStreamSynchronize(stream));  // read by kernels on the per-context stage streams
    wimg_w = w; wimg_h = h;
    return M2D_OK;
}

// Grow the pool until at least n tiles are free (128 MiB slabs).  Called from prepare() for the prepared grid, so
// that streaming feed() calls do not hit a cudaMalloc (a multi-millisecond stall) in steady state.
int m2d_map::reserve_tiles(size_t n) {
    for (;;) {
        { std::lock_guard<std::mutex> lk(pool_mu); if (free_tiles.size() >= n) return M2D_OK; }
        size_t per_chunk = std::max<size_t>(16, ((size_t)128 << 20) / tile_bytes);
        void* c = nullptr;
        cudaError_t e = cudaMalloc(&c, per_chunk * tile_bytes);
        if (e != cudaSuccess) {
            err = std::string("tile pool cudaMalloc failed: ") + cudaGetErrorString(e);
            cudaGetLastError();
            return M2D_ERR_NOMEM;
        }
        std::lock_guard<std::mutex> lk(pool_mu);
        chunks.push_back(c);
        for (size_t i = per_chunk; i-- > 0;) free_tiles.push_back((uint8_t*)c + i * tile_bytes);
    }
}

void m2d_map::pool_filler() {
    cudaSetDevice(cfg.device);
    std::unique_lock<std::mutex> lk(pool_mu);
    for (;;) {
        pool_cv.wait(lk, [&] { return pool_stop || (!pool_failed && free_tiles.size() < pool_low); });
        if (pool_stop) return;
        // a cudaMalloc stalls concurrent launches of the feed thread for a millisecond or two whatever its size: once the map
        // has grown past the first slabs, grow in 512 MiB steps so that fewer than 1 % of streamed frames ever meet one
        const size_t per_chunk = std::max<size_t>(16, ((size_t)(chunks.size() >= 8 ? 512 : 128) << 20) / tile_bytes);
        lk.unlock();
        void* c = nullptr;
        cudaError_t e = cudaMalloc(&c, per_chunk * tile_bytes);   // the slow part, off the feed path
        lk.lock();
        if (e != cudaSuccess) { cudaGetLastError(); pool_failed = true; continue; }   // the feed path will report NOMEM if it runs dry
        chunks.push_back(c);
        for (size_t i = per_chunk; i-- > 0;) free_tiles.push_back((uint8_t*)c + i * tile_bytes);
    }
}

int m2d_map::alloc_tile(uint8_t** out) {
    {
        std::lock_guard<std::mutex> lk(pool_mu);
        if (!free_tiles.empty()) {
            *out = free_tiles.back();
            free_tiles.pop_back();
            tiles_in_use++;
            if (free_tiles.size() < pool_low) pool_cv.notify_one();
            return M2D_OK;
        }
    }
    int rc = reserve_tiles(1);   // the filler fell behind (a large batch): allocate here
    if (rc != M2D_OK) return rc;
    std::lock_guard<std::mutex> lk(pool_mu);
    *out = free_tiles.back();
    free_tiles.pop_back();
    tiles_in_use++;
    return M2D_OK;
}

// (Re)allocate a device or pinned buffer.  Safe against in-flight use: callers only grow buffers of a context whose
// previous group has completed (cudaEventSynchronize on ctx.done).
int m2d_map::grow(void** p, size_t* cap, size_t need, bool pinned) {
    if (need <= *cap) return M2D_OK;
    size_t ncap = need + need / 4 + 4096;
    if (*p) { if (pinned) CU(cudaFreeHost(*p)); else CU(cudaFree(*p)); *p = nullptr; *cap = 0; }
    if (pinned) CU(cudaMallocHost(p, ncap)); else CU(cudaMalloc(p, ncap));
    *cap = ncap;
    return M2D_OK;
}

int m2d_map::group_size(int w, int h, bool on_device) const {
    if (cfg.batch_frames > 0) return std::min(cfg.batch_frames, kMaxGroup);
    // Weighted mode keeps no per-frame scratch, and its best-first culling gets sharper the more frames a group holds:
    // device-resident batches are fused in groups of 192 frames (larger single launches stop overlapping host
    // preparation); host batches stay small so that the H2D staging of one group overlaps the fusion of the previous one.
    if (type != M2D_TYPE_MULTIBAND && on_device) return 192;
    const double mpx = (double)w * h / 1e6;
    if (type == M2D_TYPE_MULTIBAND && weights_first) {
        // weights-first multi-band: the more frames compete inside a group, the smaller each frame's share of competitive
        // cells and winners, and with it the weight AND image work -> as many frames as the scratch budget holds (a frame's
        // scratch pyramid is ~11 B per px of its tile-aligned region).  Host batches keep 64-frame groups: they are bound
        // by the PCIe copy, which the next group's copy must overlap.
        if (!on_device) return std::max(4, std::min((int)(100.0 / std::max(mpx, 0.25)), 64));
        const double s = (cfg.resolution > 0 || cfg.scale <= 0) ? 1.0 : cfg.scale;
        const double tiles = (std::ceil(w * s / kEle) + 2) * (std::ceil(h * s / kEle) + 2);
        const double per_frame = tiles * kEle * kEle * 11.0;
        return std::max(4, std::min((int)(scratch_budget / per_frame), kMaxGroup));
    }
    int k = (int)(100.0 / std::max(mpx, 0.25));  // ~100 Mpx of source per group (measured: larger groups amortise better)
    return std::max(1, std::min(k, 64));
}

int m2d_map::queue_size() {
    if (!subs.empty()) { int q = 0; for (m2d_map* sub : subs) q = std::max(q, sub->queue_size()); return q; }
    int q = 0;
    for (int i = 0; i < kCtx; i++) {
        GroupCtx& c = ctx[i];
        if (c.busy && cudaEventQuery(c.done) == cudaSuccess) c.busy = false;
        if (c.busy) q += c.frames;
    }
    cudaGetLastError();
    return q;
}

int m2d_map::sync() {
    if (!subs.empty()) { int rc = M2D_OK; for (m2d_map* sub : subs) { int r = sub->sync(); if (r != M2D_OK) { rc = r; err = sub->err; } } return rc; }
    CU(cudaSetDevice(cfg.device));
    CU(cudaStreamSynchronize(stream));
    for (int i = 0; i < kCtx; i++) ctx[i].busy = false;
    return M2D_OK;
}

int m2d_map::reset() {
    if (!subs.empty()) { int rc = M2D_OK; for (m2d_map* sub : subs) { int r = sub->reset(); if (r != M2D_OK) { rc = r; err = sub->err; } } mirror(); return rc; }
    CU(cudaSetDevice(cfg.device));
    // No device synchronisation: the tiles go back to the pool while the kernels that last touched them may still be in
    // flight, and that is safe because whoever gets a recycled tile (a later group, as a FRESH tile) only ever touches it
    // from kernels that are stream-ordered behind those (decide chain -> image chain -> Laplacian chain of the handle;
    // weight planes and Laplacian planes are disjoint byte ranges of a tile).  So a reset()+feed_batch() sequence keeps
    // the GPU busy across the reset: the host prepares the next group while the previous one is still being fused.
    // With counters or per-kernel timing on, the device counters are cleared here, so everything must have landed.
    if (cfg.collect_stats || profiling) {
        CU(cudaStreamSynchronize(stream));
        for (int i = 0; i < kCtx; i++) ctx[i].busy = false;
    }
    {
        std::lock_guard<std::mutex> lk(pool_mu);
        for (uint8_t*& t : table)
            if (t) { free_tiles.push_back(t); t = nullptr; tiles_in_use--; }
    }
    std::fill(changed.begin(), changed.end(), 0);
    CU(cudaMemsetAsync(d_stats, 0, 32 * sizeof(unsigned long long), stream));
    memset(&stats, 0, sizeof stats);
    last_rect[0] = last_rect[1] = last_rect[2] = last_rect[3] = -1;
    rnd_have = false;
    return M2D_OK;
}

int m2d_map::feed_frames(int n, const uint8_t* base, size_t frame_stride, int w, int h, size_t stride, const double* poses,
                         bool on_device, int* result, const uint8_t* const* ptrs) {
    if (!subs.empty()) {
        // every sub-handle sees every frame: it takes the same accept / spreadMap decisions as an unsharded map, copies (or
        // reads over NVLink) only the frames under which it owns tiles, and returns once its work is enqueued -- the
        // devices then fuse concurrently
        std::vector<int> scratch_res((size_t)std::max(n, 1));
        int rc0 = M2D_OK;
        for (size_t i = 0; i < subs.size(); i++) {
            int r = subs[i]->feed_frames(n, base, frame_stride, w, h, stride, poses, on_device, i == 0 ? result : scratch_res.data(), ptrs);
            if (r < 0) { err = subs[i]->err; return r; }
            if (i == 0) rc0 = r;
        }
        mirror();
        return rc0;
    }
    if (!valid || type == M2D_TYPE_RENDER) {              // Map2DCPU.cpp:129; Map2DRender::renderFrame is `return false` (:464-467)
        stats.frames_fed += n;
        for (int i = 0; i < n && result; i++) result[i] = M2D_REJECTED;
        return M2D_REJECTED;
    }
    if ((!base && !ptrs) || !poses) return M2D_ERR_ARG;
    if (w != g.cam_w || h != g.cam_h) {                   // Map2DCPU.cpp:158-162
        fprintf(stderr, "Map2DB200::renderFrame: frame size != camera size\n");
        stats.frames_fed += n;
        for (int i = 0; i < n && result; i++) result[i] = M2D_REJECTED;
        return M2D_REJECTED;
    }
    if (stride < (size_t)w * 3) return M2D_ERR_ARG;
    CU(cudaSetDevice(cfg.device));
    { int rc = ensure_weight_images(w, h); if (rc != M2D_OK) return rc; }
    struct EventScope { cudaEvent_t& e; ~EventScope() { e = nullptr; } } event_scope{input_event};   // one feed call consumes the event
    // Pinned (page-locked, device-visible) host frames in batches: no whole-frame staging copies.
    const uint8_t* pull_base = nullptr;
    if (!on_device && zero_copy && !ptrs && n >= 8) {
        cudaPointerAttributes a0{}, a1{};
        const uint8_t* last = base + (size_t)(n - 1) * frame_stride + (size_t)(h - 1) * stride + (size_t)w * 3 - 1;
        if (cudaPointerGetAttributes(&a0, base) == cudaSuccess && cudaPointerGetAttributes(&a1, last) == cudaSuccess &&
            a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost && a0.devicePointer && a1.devicePointer &&
            (const uint8_t*)a1.devicePointer - (const uint8_t*)a0.devicePointer == last - base)
            pull_base = (const uint8_t*)a0.devicePointer;
        else cudaGetLastError();
    }
    int K = group_size(w, h, on_device || pull_base != nullptr);
    if (n > K) K = (n + (n + K - 1) / K - 1) / ((n + K - 1) / K);   // equal-sized groups: 500 frames at K = 480 -> 250 + 250
    int worst = M2D_OK;
    for (int i = 0; i < n; i += K) {
        int m = std::min(K, n - i);
        int rc = run_group(m, ptrs ? nullptr : base + (size_t)i * frame_stride, frame_stride, w, h, stride, poses + 7 * (size_t)i,
                           on_device, result ? result + i : nullptr, ptrs ? ptrs + i : nullptr,
                           pull_base ? pull_base + (size_t)i * frame_stride : nullptr);
        if (rc < 0) return rc;
        if (rc != M2D_OK) worst = rc;
    }
    return worst;
}

// One group: feed() semantics for frames [0,n) in order (Map2DCPU.cpp:127-336 / MultiBandMap2DCPU.cpp:288-558).
int m2d_map::run_group(int n, const uint8_t* base, size_t frame_stride, int w, int h, size_t stride, const double* poses,
                       bool on_device, int* result, const uint8_t* const* ptrs, const uint8_t* pull_base) {
    GroupCtx& c = ctx[ctx_next];
    ctx_next = (ctx_next + 1) % kCtx;
    if (c.busy) { CU(cudaEventSynchronize(c.done)); c.busy = false; }
    // Tiles this group allocates are only initialised by its kernels (TileWork::fresh).  If the group is abandoned on an
    // error before those kernels are enqueued, the slots go back to the pool: a later feed must not find them in the
    // table and blend against recycled memory.  Keyed on ABSOLUTE tile coordinates (spreadMap re-indexes the table).
    struct FreshGuard {
        m2d_map* m; std::vector<std::pair<int, int>> abs; bool armed = true;
        ~FreshGuard() {
            if (!armed) return;
            std::lock_guard<std::mutex> lk(m->pool_mu);
            for (auto& a : abs) {
                int x = a.first - m->org_x, y = a.second - m->org_y;
                if (x < 0 || y < 0 || x >= m->g.w || y >= m->g.h) continue;
                uint8_t*& slot = m->table[(size_t)y * m->g.w + x];
                if (slot) { m->free_tiles.push_back(slot); slot = nullptr; m->tiles_in_use--; }
            }
        }
    } fresh_guard{this};

    std::vector<FrameJob> jobs;
    std::vector<TileWork> tiles;
    std::vector<std::vector<TileEntry>>& per_tile = per_tile_pool;  // entry i is live for i < tiles.size()
    std::vector<std::pair<int, int>> tile_abs;  // absolute tile coordinate of every TileWork
    if (++group_epoch == 0) { std::fill(slot_epoch.begin(), slot_epoch.end(), 0u); group_epoch = 1; }
    std::vector<int> src_index;  // frame index (within the call) of every accepted job
    std::vector<std::pair<float, float>> job_centre;  // frame footprint centre in ABSOLUTE tile units (ordering heuristic)
    jobs.reserve(n);
    size_t scratch = 0;
    int max_wnx = 1, max_wny = 1, any_rejected = 0;
    const size_t npx = (size_t)w * h;

    for (int i = 0; i < n; i++) {
        stats.frames_fed++;
        int status = M2D_REJECTED;
        do {
            FrameBounds fb;
            const double* pose = poses + 7 * (size_t)i;
            frame_bounds(g, pose, &fb);
            if (!fb.ok) break;                                   // oblique view (Map2DCPU.cpp:179-182)
            if (fb.gx0 < g.min_x || fb.gx1 > g.max_x || fb.gy0 < g.min_y || fb.gy1 > g.max_y) {
                if (spread(fb.gx0, fb.gy0, fb.gx1, fb.gy1) != M2D_OK) break;   // Map2DCPU.cpp:199-218
                frame_bounds(g, pose, &fb);
                if (!fb.ok) break;
            }
            if (fb.x0 < 0 || fb.y0 < 0 || fb.x1 > g.w || fb.y1 > g.h || fb.x0 >= fb.x1 || fb.y0 >= fb.y1) {
                fprintf(stderr, "Map2DB200::renderFrame:should never happen!\n");  // Map2DCPU.cpp:223-227
                break;
            }
            int nx = fb.x1 - fb.x0, ny = fb.y1 - fb.y0;
            if (nx > 32767 || ny > 32767) { err = "frame region too large"; return M2D_ERR_UNSUPPORTED; }
            memcpy(last_rect, &fb.x0, sizeof(int) * 4);
            stats.frames_fused++;
            stats.input_px += (uint64_t)npx;
            status = M2D_OK;

            // owned tiles under the frame; first-touch allocation (Map2DCPU.cpp:311-319)
            int job_idx = (int)jobs.size();
            int ox0 = INT32_MAX, oy0 = INT32_MAX, ox1 = INT32_MIN, oy1 = INT32_MIN;
            for (int ty = fb.y0; ty < fb.y1; ty++)
                for (int tx = fb.x0; tx < fb.x1; tx++) {
                    if (!owns(tx, ty)) continue;
                    ox0 = std::min(ox0, tx); oy0 = std::min(oy0, ty); ox1 = std::max(ox1, tx + 1); oy1 = std::max(oy1, ty + 1);
                    uint8_t*& slot = table[(size_t)ty * g.w + tx];
                    bool is_fresh = false;
                    if (!slot) {
                        int rc = alloc_tile(&slot);
                        if (rc != M2D_OK) return rc;
                        is_fresh = true;
                        fresh_guard.abs.emplace_back(tx + org_x, ty + org_y);
                    }
                    const size_t gi = (size_t)ty * g.w + tx;
                    changed[gi] = 1;  // ele->Ischanged = true (Map2DCPU.cpp:330)
                    int ti;
                    if (slot_epoch[gi] != group_epoch) {
                        ti = (int)tiles.size();
                        slot_epoch[gi] = group_epoch;
                        slot_work[gi] = ti;
                        tiles.push_back(TileWork{slot, 0, 0, is_fresh ? 1 : 0});
                        if (per_tile.size() <= (size_t)ti) per_tile.emplace_back();
                        per_tile[ti].clear();
                        tile_abs.emplace_back(tx + org_x, ty + org_y);
                    } else ti = slot_work[gi];
                    per_tile[ti].push_back(TileEntry{job_idx, (short)(tx - fb.x0), (short)(ty - fb.y0)});
                    for (int l = 0; l < levels; l++) {
                        uint64_t lpx = (uint64_t)(kEle >> l) * (kEle >> l);
                        stats.region_px[l] += lpx;
                        if (is_fresh) stats.fresh_px[l] += lpx;
                    }
                }
            if (ox1 <= ox0) break;  // this shard owns nothing under the frame: accepted, no work
            FrameJob J{};
            memcpy(J.hinv, fb.hinv, sizeof J.hinv);
            for (int k = 0; k < 9; k++) J.hinvf[k] = (float)fb.hinv[k];
            J.nx = nx; J.ny = ny;
            // Pyramid window: owned tiles + a ring of tiles that covers the level-0 support of the deepest Laplacian tap
            // (G_{L-1} at +-1 coarse px of a level L-2 px: 3 * 2^(L-1) - 2 px = 94 for 6 levels -> 1 tile, 382 for 8
            // levels -> 2 tiles, 766 for 9 -> 3), clipped to the frame region so the reflect-101 border lands where the
            // reference puts it.
            if (cfg.shard_count <= 1) { J.wx = 0; J.wy = 0; J.wnx = nx; J.wny = ny; }
            else {
                const int ring = (3 * (1 << (levels - 1)) - 2 + kEle - 1) / kEle;
                int wx0 = std::max(ox0 - ring, fb.x0), wy0 = std::max(oy0 - ring, fb.y0);
                int wx1 = std::min(ox1 + ring, fb.x1), wy1 = std::min(oy1 + ring, fb.y1);
                J.wx = wx0 - fb.x0; J.wy = wy0 - fb.y0; J.wnx = wx1 - wx0; J.wny = wy1 - wy0;
            }
            max_wnx = std::max(max_wnx, J.wnx); max_wny = std::max(max_wny, J.wny);
            if (type == M2D_TYPE_MULTIBAND)
                for (int l = 0; l < levels; l++) {
                    size_t lpx = (size_t)J.wnx * J.wny * (kEle >> l) * (kEle >> l);
                    J.g_off[l] = scratch; scratch += (lpx * 4 + 255) & ~(size_t)255;
                    J.w_off[l] = scratch; scratch += (lpx * 4 + 255) & ~(size_t)255;
                }
            jobs.push_back(J);
            src_index.push_back(i);
            job_centre.emplace_back((float)((0.5 * (fb.gx0 + fb.gx1) - g.min_x) * g.ele_size_inv) + (float)org_x,
                                    (float)((0.5 * (fb.gy0 + fb.gy1) - g.min_y) * g.ele_size_inv) + (float)org_y);
        } while (0);
        if (result) result[i] = status;
        if (status != M2D_OK) any_rejected = 1;
    }
    int nj = (int)jobs.size();
    if (nj == 0) { fresh_guard.armed = false; return any_rejected ? M2D_REJECTED : M2D_OK; }

    // ---- device buffers of this context
    size_t n_entries = 0;
    for (size_t t = 0; t < tiles.size(); t++) n_entries += per_tile[t].size();
    size_t off_tiles = ((size_t)nj * sizeof(FrameJob) + 255) & ~(size_t)255;
    size_t off_entries = (off_tiles + tiles.size() * sizeof(TileWork) + 255) & ~(size_t)255;
    size_t blob = off_entries + n_entries * sizeof(TileEntry) + 256;
    { size_t cap = c.blob_cap; int rc = grow((void**)&c.h_blob, &cap, blob, true); if (rc != M2D_OK) return rc;
      rc = grow((void**)&c.d_blob, &c.blob_cap, blob, false); if (rc != M2D_OK) return rc; }
    const size_t slot_bytes = (npx * 3 + 255) & ~(size_t)255;   // staging slot of one host frame (256-byte aligned: pull mode copies 256-byte chunks)
    // Pinned host frames (pull_base = their device-visible address): not staged whole where frames overlap.  A px that a single
    // frame covers is read once either way, and SM-initiated PCIe reads are slower (38 GB/s measured) than the copy engine
    // (48-52 GB/s), so the zero-copy paths are taken only when a touched tile is covered by >= 2.5 frames on average.
    size_t region_tiles = 0;
    for (const FrameJob& J : jobs) region_tiles += (size_t)J.wnx * J.wny;
    const bool overlapping = !tiles.empty() && 2 * region_tiles >= 5 * tiles.size();
    // weighted mode: best-first culling touches every map px once or twice -> the kernel samples the host frames in place
    const bool inplace = !on_device && pull_base && !ptrs && overlapping && type != M2D_TYPE_MULTIBAND;
    if (!on_device && !inplace) { int rc = grow((void**)&c.d_raw, &c.raw_cap, (size_t)nj * slot_bytes + 256, false); if (rc != M2D_OK) return rc; }
    // weights-first multi-band: cell flags, competitive masks, work lists and the winner map live behind the pyramids
    // (groups of a few frames -- streaming feed() calls -- take the dense pipeline: a lone frame wins most of what it
    // covers, so there is little to skip, and the dense pipeline needs 6 launches instead of ~20)
    const bool sparse = weights_first && type == M2D_TYPE_MULTIBAND && levels <= 6 && !tiles.empty() && nj >= 4;
    const int cells_max = max_wnx * 8 * max_wny * 8;
    const int wmap_stride = (lay.px_off[levels] + 7) & ~7;
    size_t off_flags = 0, off_cmask = 0, off_lists = 0, off_counts = 0, off_wmap = 0, off_etable = 0, flag_bytes = 0, off_srcbits = 0;
    int mask_words = 1;
    // pull mode: pinned host frames, tightly packed, 16-byte aligned -> the image stage pulls the chunks it needs itself
    const bool pull = sparse && !on_device && pull_base && !ptrs && overlapping && stride == (size_t)w * 3 &&
                      ((uintptr_t)pull_base % 16 == 0) && (frame_stride % 16 == 0);
    const int src_words = (int)(((npx * 3 + 255) / 256 + 31) / 32);
    if (sparse) {
        if (cells_max > 65535 || nj > 65535) { err = "frame region too large for the weights-first work lists"; return M2D_ERR_UNSUPPORTED; }
        size_t max_count = 1;
        for (size_t t = 0; t < tiles.size(); t++) max_count = std::max(max_count, per_tile[t].size());
        mask_words = (int)((max_count + 31) / 32);
        flag_bytes = ((size_t)nj * levels * cells_max + 255) & ~(size_t)255;
        off_flags = scratch; scratch += 2 * flag_bytes;                                   // comp | win   (zeroed per group ...
        off_counts = scratch; scratch += 256;                                             // ... together with the 12 list lengths)
        scratch += 2 * flag_bytes;                                                        // needw | need (fully written by propagate)
        off_cmask = scratch; scratch += ((size_t)tiles.size() * levels * 64 * mask_words * sizeof(uint32_t) + 255) & ~(size_t)255;
        off_lists = scratch; scratch += ((size_t)2 * levels * nj * cells_max * sizeof(uint32_t) + 255) & ~(size_t)255;
        off_wmap = scratch; scratch += ((size_t)tiles.size() * wmap_stride * sizeof(uint16_t) + 255) & ~(size_t)255;
        off_etable = scratch; scratch += (n_entries * levels * sizeof(EntryRef) + 255) & ~(size_t)255;
        if (pull) { off_srcbits = scratch; scratch += ((size_t)nj * src_words * sizeof(uint32_t) + 255) & ~(size_t)255; }
        if ((size_t)nj * levels * cells_max > 0x7fffffffull) { err = "group too large for 32-bit cell flag indices"; return M2D_ERR_UNSUPPORTED; }
    }
    if (scratch) { int rc = grow((void**)&c.d_scratch, &c.scratch_cap, scratch, false); if (rc != M2D_OK) return rc; }

    // ---- frames: host images are staged into HBM on the copy stream (this context's previous group has finished,
    // so d_raw is free; the copy overlaps the OTHER context's kernels); device images are used in place
    const bool tight = !ptrs && stride == (size_t)w * 3 && frame_stride == npx * 3;  // frames back to back: copy runs, not frames
    for (int j = 0; j < nj; j++) {
        const uint8_t* src = ptrs ? ptrs[src_index[j]] : base + (size_t)src_index[j] * frame_stride;
        if (on_device) { jobs[j].raw = src; jobs[j].raw_stride = (int)stride; }
        else if (inplace) { jobs[j].raw = pull_base + (size_t)src_index[j] * frame_stride; jobs[j].raw_stride = (int)stride; }
        else if (pull) {
            jobs[j].raw = c.d_raw + (size_t)j * slot_bytes; jobs[j].raw_stride = w * 3;
            jobs[j].pull_src = pull_base + (size_t)src_index[j] * frame_stride;
        } else {
            uint8_t* dst = c.d_raw + (size_t)j * slot_bytes;
            if (tight && slot_bytes == npx * 3) {
                if (j == 0 || src_index[j] != src_index[j - 1] + 1) {  // start of a run of consecutive accepted frames
                    int e = j;
                    while (e + 1 < nj && src_index[e + 1] == src_index[e] + 1) e++;
                    CU(cudaMemcpyAsync(dst, src, (size_t)(e - j + 1) * npx * 3, cudaMemcpyHostToDevice, copy_stream));
                }
            } else if (stride == (size_t)w * 3) CU(cudaMemcpyAsync(dst, src, npx * 3, cudaMemcpyHostToDevice, copy_stream));
            else CU(cudaMemcpy2DAsync(dst, (size_t)w * 3, src, stride, (size_t)w * 3, h, cudaMemcpyHostToDevice, copy_stream));
            jobs[j].raw = dst; jobs[j].raw_stride = w * 3;
        }
    }
    if (!on_device) {
        CU(cudaEventRecord(c.copied, copy_stream));
        if (!sparse) CU(cudaStreamWaitEvent(c.stage, c.copied, 0));   // weights-first: only the image stage waits for the pixels
    }
    // ---- work lists
    memcpy(c.h_blob, jobs.data(), (size_t)nj * sizeof(FrameJob));
    TileEntry* he = reinterpret_cast<TileEntry*>(c.h_blob + off_entries);
    int first = 0;
    // Weighted mode visits a tile's frames best-first (closest footprint centre first): the result is order-free
    // ("largest alpha, earliest frame on ties", tracked per px by the kernel) and almost every later frame is then
    // rejected by the alpha upper bound before any sampling.  With collect_stats the sequential order is kept.
    // (multi-band keeps feed order: its tie rule -- the LAST of equal weights wins -- is resolved by scanning in order)
    const bool best_first = (type != M2D_TYPE_MULTIBAND) && !cfg.collect_stats;
    for (size_t t = 0; t < tiles.size(); t++) {
        if (best_first && per_tile[t].size() > 1) {
            float cx = (float)tile_abs[t].first + 0.5f, cy = (float)tile_abs[t].second + 0.5f;
            std::stable_sort(per_tile[t].begin(), per_tile[t].end(), [&](const TileEntry& a, const TileEntry& b) {
                float ax = job_centre[a.frame].first - cx, ay = job_centre[a.frame].second - cy;
                float bx = job_centre[b.frame].first - cx, by = job_centre[b.frame].second - cy;
                return ax * ax + ay * ay < bx * bx + by * by;
            });
        }
        tiles[t].first = first;
        tiles[t].count = (int)per_tile[t].size();
        memcpy(he + first, per_tile[t].data(), per_tile[t].size() * sizeof(TileEntry));
        first += tiles[t].count;
    }
    memcpy(c.h_blob + off_tiles, tiles.data(), tiles.size() * sizeof(TileWork));
    CU(cudaMemcpyAsync(c.d_blob, c.h_blob, blob, cudaMemcpyHostToDevice, c.stage));

    GroupParams p{};
    p.jobs = reinterpret_cast<const FrameJob*>(c.d_blob);
    p.tiles = reinterpret_cast<const TileWork*>(c.d_blob + off_tiles);
    p.entries = reinterpret_cast<const TileEntry*>(c.d_blob + off_entries);
    p.n_frames = nj; p.n_tiles = (int)tiles.size();
    p.sw = w; p.sh = h; p.levels = levels; p.weight_type = cfg.weight_type; p.f32_mode = cfg.f32_mode;
    p.alpha = d_alpha; p.wimg = d_wimg; p.scratch = c.d_scratch;
    p.stats = cfg.collect_stats ? d_stats : nullptr;
    p.need_stats = (cfg.collect_stats || profiling) ? d_stats : nullptr;
    p.max_wnx = max_wnx; p.max_wny = max_wny;
    if (sparse) {
        p.comp = c.d_scratch + off_flags; p.win = p.comp + flag_bytes;
        p.needw = c.d_scratch + off_counts + 256; p.need = p.needw + flag_bytes;
        p.cells_max = cells_max;
        p.cmask = reinterpret_cast<uint32_t*>(c.d_scratch + off_cmask); p.mask_words = mask_words;
        p.lists = reinterpret_cast<uint32_t*>(c.d_scratch + off_lists); p.list_cap = nj * cells_max;
        p.list_count = reinterpret_cast<unsigned*>(c.d_scratch + off_counts);
        p.wmap = reinterpret_cast<uint16_t*>(c.d_scratch + off_wmap); p.wmap_stride = wmap_stride;
        p.etable = reinterpret_cast<EntryRef*>(c.d_scratch + off_etable);
        make_reach_table(levels, p.reach_lo, p.reach_hi);
        make_weight_reach_table(levels, p.wreach_lo, p.wreach_hi);
        p.use_tma = use_tma ? 1 : 0;
        p.cull = (weight_cull && !cfg.collect_stats) ? 1 : 0;   // the win counters follow the sequential semantics: no culling then
        if (pull) { p.src_bits = reinterpret_cast<uint32_t*>(c.d_scratch + off_srcbits); p.src_words = src_words; p.frame_bytes = npx * 3; }
    }

    // Order-independent stages run on the context's own stream (they overlap the previous group's select); the
    // order-dependent tile-centric stage runs on the handle's stream, which serialises groups in feed order.
    // While m2d_profile is on, nothing may overlap (per-kernel event timings must be clean): the stage stream
    // first waits for everything already enqueued on the handle's stream.
    if (profiling) {
        CU(cudaEventRecord(c.staged, stream));
        CU(cudaStreamWaitEvent(c.stage, c.staged, 0));
    }
    if (sparse) {
        const int ctas = sm_count * 8;   // persistent list kernels: every resident CTA slot of the GPU, grid-stride over the items
        // 0.+1. competitive cells (closed-form bounds against the other frames and the tile state), then the weights in the
        // cells that matter.  bounds reads/reset cmin, which the previous group's decide wrote: it runs on the decide chain.
        cudaStream_t ds = profiling ? stream : decide_stream;
        CU(cudaEventRecord(c.staged, c.stage));                      // the blob upload
        CU(cudaStreamWaitEvent(ds, c.staged, 0));
        if (dense_pending) {   // a dense group's select (handle stream) wrote tile weights: the decide chain must see them
            CU(cudaStreamWaitEvent(ds, dense_done, 0));
            dense_pending = false;
        }
        CU(cudaMemsetAsync(p.comp, 0, 2 * flag_bytes + 256, ds));    // comp | win flags ... and the list lengths (off_counts follows the flags)
        LAUNCHKS(M2D_K_MBC_BOUNDS, ds, launch_mbc_bounds(p, lay, ds));
        LAUNCHKS(M2D_K_MBC_BOUNDS, ds, launch_mbc_entry_table(p, (int)n_entries, ds));
        LAUNCHKS(M2D_K_MBS_PROPAGATE, ds, launch_mbx_propagate(p, 0, ds));
        CU(cudaEventRecord(c.decided, ds));
        CU(cudaStreamWaitEvent(c.stage, c.decided, 0));
        LAUNCHKS(M2D_K_MBW_WARP, c.stage, launch_mbw_warp(p, ctas, c.stage));
        for (int l = 0; l + 1 < levels; l++) LAUNCHKS(M2D_K_MBW_PYR, c.stage, launch_mbx_pyrdown(p, 0, l, ctas, c.stage));
        CU(cudaEventRecord(c.staged, c.stage));
        // 2.+3. winners and need flags: the decide chain serialises groups in feed order (tile weights only), and runs
        // ahead of the Laplacian chain on the handle's stream (which writes the Laplacian planes only)
        CU(cudaStreamWaitEvent(ds, c.staged, 0));
        LAUNCHKS(M2D_K_MBS_DECIDE, ds, launch_mbs_decide(p, lay, ds));
        LAUNCHKS(M2D_K_MBS_PROPAGATE, ds, launch_mbx_propagate(p, 1, ds));
        CU(cudaEventRecord(c.decided, ds));
        // 4. image work in the needed cells, back on the context's stream (only here are the frames' pixels needed)
        CU(cudaStreamWaitEvent(c.stage, c.decided, 0));
        if (!on_device && !pull) CU(cudaStreamWaitEvent(c.stage, c.copied, 0));
        if (input_event) CU(cudaStreamWaitEvent(c.stage, input_event, 0));   // frames still travelling (e.g. NCCL halo exchange)
        if (pull) {   // the GPU fetches the chunks of the pinned host frames that the listed cells sample -- nothing else crosses PCIe
            CU(cudaMemsetAsync(p.src_bits, 0, (size_t)nj * src_words * sizeof(uint32_t), c.stage));
            if (pull_poison) CU(cudaMemsetAsync(c.d_raw, 0xA5, (size_t)nj * slot_bytes, c.stage));
            LAUNCHKS(M2D_K_MISC, c.stage, launch_mbs_mark(p, ctas, c.stage));
            LAUNCHKS(M2D_K_MISC, c.stage, launch_mbs_pull(p, ctas, c.stage));
        }
        LAUNCHKS(M2D_K_MBS_WARP, c.stage, launch_mbs_warp(p, ctas, c.stage));
        for (int l = 0; l + 1 < levels; l++) LAUNCHKS(M2D_K_MBS_PYR, c.stage, launch_mbx_pyrdown(p, 1, l, ctas, c.stage));
        CU(cudaEventRecord(c.staged, c.stage));
        // 5. winners' Laplacians into the tiles, in feed order on the handle's stream
        CU(cudaStreamWaitEvent(stream, c.staged, 0));
        LAUNCHK(M2D_K_MBS_LAP, launch_mbs_lap(p, lay, stream));
    } else if (type == M2D_TYPE_MULTIBAND) {
        // full-grid pyrDown while a level is big enough; the small deep levels go through one tail launch
        int l = 0;
        if (input_event) CU(cudaStreamWaitEvent(c.stage, input_event, 0));
        LAUNCHKS(M2D_K_MB_WARP, c.stage, launch_mb_warp(p, c.stage));
        // (a level goes to the tail only when its output is small: <= 16 K px per frame, i.e. one 1024-thread CTA's worth)
        auto small_level = [&](int lv) { long long nd = kEle >> (lv + 1); return (long long)max_wnx * nd * max_wny * nd <= 16384; };
        for (; l + 1 < levels && (!small_level(l) || levels - 1 - l < 2); l++) LAUNCHKS(M2D_K_MB_PYRDOWN, c.stage, launch_mb_pyrdown(p, l, c.stage));
        if (l + 1 < levels) LAUNCHKS(M2D_K_MB_PYRTAIL, c.stage, launch_mb_pyrtail(p, l, c.stage));
        CU(cudaEventRecord(c.staged, c.stage));
        CU(cudaStreamWaitEvent(stream, c.staged, 0));
        LAUNCHK(M2D_K_MB_SELECT, launch_mb_select(p, lay, stream));
        if (weights_first) { CU(cudaEventRecord(dense_done, stream)); dense_pending = true; }
    } else {
        // weighted mode samples the caller's BGR8 frames in place (plus the alpha plane): no packed copy
        if (input_event) CU(cudaStreamWaitEvent(c.stage, input_event, 0));
        CU(cudaEventRecord(c.staged, c.stage));
        CU(cudaStreamWaitEvent(stream, c.staged, 0));
        LAUNCHK(M2D_K_WEIGHTED, launch_weighted_group(p, stream));
    }
    fresh_guard.armed = false;   // every kernel of the group is enqueued: the fresh tiles will be initialised
    CU(cudaEventRecord(c.done, stream));
    c.busy = true;
    c.frames = nj;
    return any_rejected ? M2D_REJECTED : M2D_OK;
}

bool m2d_map::tile_bbox(int& x0, int& y0, int& x1, int& y1) const {
    x0 = y0 = INT32_MAX; x1 = y1 = INT32_MIN;
    for (int y = 0; y < g.h; y++)
        for (int x = 0; x < g.w; x++)
            if (table[(size_t)y * g.w + x]) {
                x0 = std::min(x0, x); y0 = std::min(y0, y); x1 = std::max(x1, x + 1); y1 = std::max(y1, y + 1);
            }
    return x1 > x0;
}

// save() in memory — Map2DCPU.cpp:523-560 / MultiBandMap2DCPU.cpp:779-841
int m2d_map::get_image(uint8_t* out, int* w, int* h, int* channels, int* tmx, int* tmy) {
    if (!subs.empty()) return multi_get_image(out, w, h, channels, tmx, tmy);
    if (!valid || g.w == 0 || g.h == 0) return M2D_REJECTED;
    if (type == M2D_TYPE_RENDER) {   // the blended canvas of the last m2d_render_frames, 16S -> 8U like cv::imwrite / convertTo (:748-749)
        if (!rnd_have) return M2D_REJECTED;
        *w = rnd_w; *h = rnd_h; *channels = 3; *tmx = rnd_tx0 - org_x; *tmy = rnd_ty0 - org_y;
        if (!out) return M2D_OK;
        CU(cudaSetDevice(cfg.device));
        CU(cudaMemcpyAsync(out, rnd_canvas + rnd_off8, (size_t)rnd_w * rnd_h * 3, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return M2D_OK;
    }
    int x0, y0, x1, y1;
    if (!tile_bbox(x0, y0, x1, y1)) return M2D_REJECTED;
    *tmx = x0; *tmy = y0;
    const int win[4] = {x0, y0, x1, y1};
    return collapse_window(out, false, win, win, w, h, channels);
}

// The collapse of save() over an explicit WINDOW of tiles (grid coordinates; tiles the table does not hold count as the
// zeros the reference pastes for absent tiles), of which the CROP rows/columns are written to `out` (host or device).
// The whole-map save is window = crop = bbox of the touched tiles.  A sharded save collapses, on every shard, the global
// bbox's columns x (its own tile rows + halo rows) and crops its own rows: a px of restored level l-1 depends on
// restored level l within 1 coarse px, so what a window edge that is NOT the mosaic's edge gets wrong reaches
// 2^levels - 2 level-0 px into the window (62 px for 6 levels) -- one halo tile row keeps the crop exact.
int m2d_map::collapse_window(uint8_t* out, bool out_on_device, const int win[4], const int crop[4], int* w, int* h, int* channels) {
    if (!valid) return M2D_REJECTED;
    const int x0 = win[0], y0 = win[1], x1 = win[2], y1 = win[3];
    if (x1 <= x0 || y1 <= y0 || crop[0] < x0 || crop[1] < y0 || crop[2] > x1 || crop[3] > y1 || crop[2] <= crop[0] || crop[3] <= crop[1]) return M2D_ERR_ARG;
    int tw = x1 - x0, th = y1 - y0;
    int cn = (type == M2D_TYPE_MULTIBAND) ? 3 : 4;
    const size_t CW = (size_t)(crop[2] - crop[0]) * kEle, CH = (size_t)(crop[3] - crop[1]) * kEle;
    *w = (int)CW; *h = (int)CH; *channels = cn;
    if (!out) return M2D_OK;
    CU(cudaSetDevice(cfg.device));
    size_t W = (size_t)tw * kEle, H = (size_t)th * kEle;
    const cudaMemcpyKind kind = out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const bool full = crop[0] == x0 && crop[1] == y0 && crop[2] == x1 && crop[3] == y1;
    auto tile_at = [&](int x, int y) -> const uint8_t* {
        return (x < 0 || y < 0 || x >= g.w || y >= g.h) ? nullptr : table[(size_t)y * g.w + x];
    };
    std::vector<PasteItem> items;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++)
            if (const uint8_t* t = tile_at(x, y)) items.push_back(PasteItem{t, x - x0, y - y0});
    if (type != M2D_TYPE_MULTIBAND) {
        // assemble the BGRA mosaic in HBM (one paste launch), then ONE copy out
        size_t off_items = (W * H * 4 + 255) & ~(size_t)255;
        CU(cudaStreamSynchronize(stream));
        { int rc = grow((void**)&d_collapse, &collapse_cap, off_items + items.size() * sizeof(PasteItem) + 256, false); if (rc != M2D_OK) return rc; }
        PasteItem* d_items = reinterpret_cast<PasteItem*>(d_collapse + off_items);
        CU(cudaMemsetAsync(d_collapse, 0, W * H * 4, stream));  // untouched tiles: the reference leaves them undefined, we define 0
        CU(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(PasteItem), cudaMemcpyHostToDevice, stream));
        LAUNCHK(M2D_K_COLLAPSE, launch_bgra_paste(d_items, (int)items.size(), reinterpret_cast<uint32_t*>(d_collapse), (int)W, stream));
        if (full) CU(cudaMemcpyAsync(out, d_collapse, W * H * 4, kind, stream));
        else CU(cudaMemcpy2DAsync(out, CW * 4, d_collapse + ((size_t)(crop[1] - y0) * kEle * W + (size_t)(crop[0] - x0) * kEle) * 4, W * 4, CW * 4, CH, kind, stream));
        CU(cudaStreamSynchronize(stream));
        return M2D_OK;
    }
    // multi-band: collapse on the GPU.  Buffers are cached in the handle (grow-only): cudaMalloc/cudaFree of ~1 GB
    // per call costs far more than the collapse itself.
    MosaicSet ms{};
    size_t off = 0, offs[M2D_MAX_LEVELS][3];
    for (int l = 0; l < levels; l++) {
        int n = kEle >> l;
        ms.lv[l].w = tw * n; ms.lv[l].h = th * n;
        size_t px = (size_t)ms.lv[l].w * ms.lv[l].h;
        for (int c = 0; c < 3; c++) { offs[l][c] = off; off += (px * sizeof(int16_t) + 255) & ~(size_t)255; }
    }
    size_t off_w0 = off; off += (W * H * sizeof(float) + 255) & ~(size_t)255;
    size_t zero_bytes = off;  // everything up to here must start at 0 (tiles never touched stay 0 / weight 0)
    size_t off_out = off; off += (W * H * 3 + 255) & ~(size_t)255;
    size_t off_items = off; off += items.size() * sizeof(PasteItem) + 256;
    CU(cudaStreamSynchronize(stream));
    { int rc = grow((void**)&d_collapse, &collapse_cap, off, false); if (rc != M2D_OK) return rc; }
    for (int l = 0; l < levels; l++)
        for (int c = 0; c < 3; c++) ms.lv[l].g[c] = reinterpret_cast<int16_t*>(d_collapse + offs[l][c]);
    ms.w0 = reinterpret_cast<float*>(d_collapse + off_w0);
    uint8_t* d_out = d_collapse + off_out;
    PasteItem* d_items = reinterpret_cast<PasteItem*>(d_collapse + off_items);
    CU(cudaMemsetAsync(d_collapse, 0, zero_bytes, stream));
    CU(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(PasteItem), cudaMemcpyHostToDevice, stream));
    LAUNCHK(M2D_K_COLLAPSE, launch_mosaic_paste(d_items, (int)items.size(), lay, ms, stream));
    for (int l = levels - 1; l > 0; l--) LAUNCHK(M2D_K_COLLAPSE, launch_mosaic_upadd(ms.lv[l], ms.lv[l - 1], stream));
    LAUNCHK(M2D_K_COLLAPSE, launch_mosaic_final(ms.lv[0], ms.w0, cfg.background, d_out, stream));
    if (full) CU(cudaMemcpyAsync(out, d_out, W * H * 3, kind, stream));
    else CU(cudaMemcpy2DAsync(out, CW * 3, d_out + ((size_t)(crop[1] - y0) * kEle * W + (size_t)(crop[0] - x0) * kEle) * 3, W * 3, CW * 3, CH, kind, stream));
    CU(cudaStreamSynchronize(stream));
    return M2D_OK;
}

// Multi-device save: the raw tiles of the other devices are read into the first device's pool over NVLink (peer loads from
// the import kernel), collapsed there like a single-GPU map, and the copies are given back.  (The strip-wise sharded save
// that never holds the whole map on one GPU is the multi-process path: pi-slam-fusion_b200/sharded.py save_sharded.)
int m2d_map::multi_get_image(uint8_t* out, int* w, int* h, int* channels, int* tmx, int* tmy) {
    int X0 = INT32_MAX, Y0 = INT32_MAX, X1 = INT32_MIN, Y1 = INT32_MIN;
    for (m2d_map* sub : subs) {
        int x0, y0, x1, y1;
        if (!sub->valid || !sub->tile_bbox(x0, y0, x1, y1)) continue;
        X0 = std::min(X0, x0); Y0 = std::min(Y0, y0); X1 = std::max(X1, x1); Y1 = std::max(Y1, y1);
    }
    if (X1 <= X0) return M2D_REJECTED;
    *w = (X1 - X0) * kEle; *h = (Y1 - Y0) * kEle; *channels = (type == M2D_TYPE_MULTIBAND) ? 3 : 4; *tmx = X0; *tmy = Y0;
    if (!out) return M2D_OK;
    m2d_map* root = subs[0];
    std::vector<int> foreign;   // absolute coordinates of the imported copies
    int rc = M2D_OK;
    for (size_t i = 1; i < subs.size() && rc == M2D_OK; i++) {
        m2d_map* sub = subs[i];
        const int n = (int)sub->tiles_in_use;
        if (!n) continue;
        CU(cudaSetDevice(sub->cfg.device));
        uint8_t* buf = nullptr;
        if (cudaMalloc(&buf, (size_t)n * sub->tile_bytes) != cudaSuccess) { cudaGetLastError(); err = "multi-device save: staging allocation failed"; return M2D_ERR_NOMEM; }
        std::vector<int> xy((size_t)n * 2);
        int got = 0;
        rc = m2d_export_tiles(sub, n, xy.data(), buf, 1, &got);
        if (rc == M2D_OK) rc = m2d_import_tiles(root, got, xy.data(), buf, 1);   // runs on the root device, reads `buf` through peer access
        if (rc != M2D_OK) err = root->err.empty() ? sub->err : root->err;
        cudaSetDevice(sub->cfg.device);
        cudaFree(buf);
        foreign.insert(foreign.end(), xy.begin(), xy.begin() + 2 * (size_t)got);
    }
    if (rc == M2D_OK) { rc = root->get_image(out, w, h, channels, tmx, tmy); if (rc != M2D_OK) err = root->err; }
    {   // the root keeps only what it owns
        cudaSetDevice(root->cfg.device);
        cudaStreamSynchronize(root->stream);
        std::lock_guard<std::mutex> lk(root->pool_mu);
        for (size_t k = 0; k + 1 < foreign.size(); k += 2) {
            const int x = foreign[k] - root->org_x, y = foreign[k + 1] - root->org_y;
            if (x < 0 || y < 0 || x >= root->g.w || y >= root->g.h) continue;
            uint8_t*& t = root->table[(size_t)y * root->g.w + x];
            if (t) { root->free_tiles.push_back(t); t = nullptr; root->tiles_in_use--; }
        }
    }
    return rc;
}

#define MULTI_UNSUPPORTED(h) do { if ((h) && !(h)->subs.empty()) { (h)->err = std::string(__func__) + ": not available on a multi-device handle"; return M2D_ERR_UNSUPPORTED; } } while (0)

// ---------------------------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

void m2d_config_default(m2d_config* c) {
    memset(c, 0, sizeof *c);
    c->scale = 1.0; c->band_number = 5; c->shard_count = 1; c->shard_axis = 1; c->shard_span = 4;
}

int m2d_create(int type, const m2d_config* cfg, m2d_handle* out) {
    if (!out) return M2D_ERR_ARG;
    *out = nullptr;
    if (type == M2D_TYPE_GPU) type = M2D_TYPE_CPU;  // Map2D.cpp:57-65: TypeGPU yields the Map2DCPU semantics
    if (type != M2D_TYPE_CPU && type != M2D_TYPE_MULTIBAND && type != M2D_TYPE_RENDER) return M2D_ERR_UNSUPPORTED;
    m2d_config c;
    if (cfg) c = *cfg; else m2d_config_default(&c);
    if (c.force_float) return M2D_ERR_UNSUPPORTED;
    if (type == M2D_TYPE_RENDER && (c.render_blend < 0 || c.render_blend > 2 || c.render_bands < 0 || c.shard_count > 1)) return M2D_ERR_ARG;
    if (c.f32_mode != 0 && c.f32_mode != 1) return M2D_ERR_ARG;
    if (c.scale == 0) c.scale = 1.0;
    if (c.shard_count < 1) c.shard_count = 1;
    if (c.shard_rank < 0 || c.shard_rank >= c.shard_count) return M2D_ERR_ARG;
    m2d_map* m = new m2d_map();
    m->type = type;
    m->cfg = c;
    int bn = c.band_number > 0 ? c.band_number : 5;
    m->band_num = std::min(bn, (int)ceil(log((double)kEle) / log(2.0)));  // MultiBandMap2DCPU.cpp:263
    int rc = m->init();
    if (rc != M2D_OK) {
        fprintf(stderr, "m2d_create: %s\n", m->err.c_str());
        m->release();
        delete m;
        return rc;
    }
    *out = m;
    return M2D_OK;
}

int m2d_create_multi(int type, const m2d_config* cfg, int n_devices, const int* devices, m2d_handle* out) {
    if (!out) return M2D_ERR_ARG;
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64 || !devices) return M2D_ERR_ARG;
    m2d_config c;
    if (cfg) c = *cfg; else m2d_config_default(&c);
    if (n_devices == 1) { c.device = devices[0]; return m2d_create(type, &c, out); }
    if (type == M2D_TYPE_RENDER) return M2D_ERR_UNSUPPORTED;   // one batch, one canvas: a single device
    if (c.shard_count > 1) return M2D_ERR_ARG;   // the multi-device handle shards by itself
    if (c.shard_axis != 0 && c.shard_axis != 1) c.shard_axis = 1;
    if (c.shard_span < 1) c.shard_span = 4;
    m2d_map* top = new m2d_map();
    top->cfg = c;
    top->cfg.device = devices[0];
    for (int i = 0; i < n_devices; i++) {
        m2d_config ci = c;
        ci.device = devices[i]; ci.shard_rank = i; ci.shard_count = n_devices;
        m2d_handle sub = nullptr;
        int rc = m2d_create(type, &ci, &sub);
        if (rc != M2D_OK) { top->release(); delete top; return rc; }
        top->subs.push_back(sub);
    }
    top->type = top->subs[0]->type; top->band_num = top->subs[0]->band_num; top->levels = top->subs[0]->levels;
    top->lay = top->subs[0]->lay; top->tile_bytes = top->subs[0]->tile_bytes;
    for (int a = 0; a < n_devices; a++)      // peer access both ways: frames and tiles are read in place over NVLink
        for (int b = 0; b < n_devices; b++) {
            if (a == b || devices[a] == devices[b]) continue;
            int can = 0;
            if (cudaSetDevice(devices[a]) == cudaSuccess && cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); can = 0; }
                cudaGetLastError();
            }
            if (!can) { fprintf(stderr, "m2d_create_multi: no peer access between devices %d and %d\n", devices[a], devices[b]); top->release(); delete top; return M2D_ERR_UNSUPPORTED; }
        }
    *out = top;
    return M2D_OK;
}

void m2d_destroy(m2d_handle h) {
    if (!h) return;
    m2d_ingest_close(h);
    h->release();
    delete h;
}

int m2d_prepare(m2d_handle h, const double* plane, const double* camera, int n, const double* poses) {
    API_LOCK(h);
    if (!h || !plane || !camera) return M2D_ERR_ARG;
    return h->prepare(plane, camera, n, poses);
}

int m2d_feed(m2d_handle h, const uint8_t* bgr, int w, int hpx, size_t stride, const double* pose) {
    API_LOCK(h);
    if (!h) return M2D_ERR_ARG;
    int r = M2D_REJECTED;
    int rc = h->feed_frames(1, bgr, 0, w, hpx, stride, pose, false, &r);
    return rc < 0 ? rc : r;
}
int m2d_feed_device(m2d_handle h, const uint8_t* d_bgr, int w, int hpx, size_t stride, const double* pose) {
    API_LOCK(h);
    if (!h) return M2D_ERR_ARG;
    int r = M2D_REJECTED;
    int rc = h->feed_frames(1, d_bgr, 0, w, hpx, stride, pose, true, &r);
    return rc < 0 ? rc : r;
}
int m2d_feed_batch(m2d_handle h, int n, const uint8_t* base, size_t frame_stride, int w, int hpx, size_t stride,
                   const double* poses, int on_device, int* result) {
    API_LOCK(h);
    if (!h || n < 0) return M2D_ERR_ARG;
    if (n == 0) return M2D_OK;  // empty batch: nothing to do
    if (!base || !poses) return M2D_ERR_ARG;
    int rc = h->feed_frames(n, base, frame_stride, w, hpx, stride, poses, on_device != 0, result);
    return rc < 0 ? rc : M2D_OK;
}

int m2d_feed_batch_ptrs(m2d_handle h, int n, const uint8_t* const* frames, int w, int hpx, size_t stride, const double* poses,
                        int on_device, int* result) {
    API_LOCK(h);
    if (!h || n < 0) return M2D_ERR_ARG;
    if (n == 0) return M2D_OK;
    if (!frames || !poses) return M2D_ERR_ARG;
    int rc = h->feed_frames(n, nullptr, 0, w, hpx, stride, poses, on_device != 0, result, frames);
    return rc < 0 ? rc : M2D_OK;
}

int m2d_feed_poses(m2d_handle h, int n, const double* poses, int* result) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || n < 0) return M2D_ERR_ARG;
    if (n == 0) return M2D_OK;
    if (!poses) return M2D_ERR_ARG;
    m2d_map& m = *h;
    int violations = 0;
    for (int i = 0; i < n; i++) {
        m.stats.frames_fed++;
        int status = M2D_REJECTED;
        do {
            if (!m.valid) break;
            FrameBounds fb;
            const double* pose = poses + 7 * (size_t)i;
            frame_bounds(m.g, pose, &fb, false);   // no pixels: the homography is not needed
            if (!fb.ok) break;
            if (fb.gx0 < m.g.min_x || fb.gx1 > m.g.max_x || fb.gy0 < m.g.min_y || fb.gy1 > m.g.max_y) {
                if (m.spread(fb.gx0, fb.gy0, fb.gx1, fb.gy1) != M2D_OK) break;
                frame_bounds(m.g, pose, &fb, false);   // no pixels: the homography is not needed
                if (!fb.ok) break;
            }
            if (fb.x0 < 0 || fb.y0 < 0 || fb.x1 > m.g.w || fb.y1 > m.g.h || fb.x0 >= fb.x1 || fb.y0 >= fb.y1) break;
            memcpy(m.last_rect, &fb.x0, sizeof(int) * 4);
            status = M2D_OK;
            for (int ty = fb.y0; ty < fb.y1 && status == M2D_OK; ty++)
                for (int tx = fb.x0; tx < fb.x1; tx++)
                    if (m.owns(tx, ty)) { status = M2D_ERR_ARG; violations++; break; }
        } while (0);
        if (result) result[i] = status;
    }
    if (violations) {
        m.err = "m2d_feed_poses: " + std::to_string(violations) + " pose(s) touch tiles this shard owns; their pixels are required";
        return M2D_ERR_ARG;
    }
    return M2D_OK;
}

int m2d_plan_rects(m2d_handle h, int n, const double* poses, int* rects) {
    API_LOCK(h);
    if (h && !h->subs.empty()) return m2d_plan_rects(h->subs[0], n, poses, rects);
    if (!h || n < 0 || (n && (!poses || !rects))) return M2D_ERR_ARG;
    if (!h->valid) return M2D_ERR_STATE;
    GridGeom g = h->g;   // dry run on a copy: the map itself is not touched
    int org_x = h->org_x, org_y = h->org_y;
    for (int i = 0; i < n; i++) {
        int* r = rects + 4 * (size_t)i;
        r[0] = r[1] = r[2] = r[3] = -1;
        FrameBounds fb;
        const double* pose = poses + 7 * (size_t)i;
        frame_bounds(g, pose, &fb, false);
        if (!fb.ok) continue;
        if (fb.gx0 < g.min_x || fb.gx1 > g.max_x || fb.gy0 < g.min_y || fb.gy1 > g.max_y) {
            int dx, dy;
            if (grow_geom(g, org_x, org_y, fb.gx0, fb.gy0, fb.gx1, fb.gy1, &dx, &dy) != M2D_OK) continue;
            frame_bounds(g, pose, &fb, false);
            if (!fb.ok) continue;
        }
        if (fb.x0 < 0 || fb.y0 < 0 || fb.x1 > g.w || fb.y1 > g.h || fb.x0 >= fb.x1 || fb.y0 >= fb.y1) continue;
        r[0] = fb.x0 + org_x; r[1] = fb.y0 + org_y; r[2] = fb.x1 + org_x; r[3] = fb.y1 + org_y;
    }
    return M2D_OK;
}

int m2d_set_shard(m2d_handle h, int rank, int count, int axis, int span, int origin) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || count < 1 || rank < 0 || rank >= count || (axis != 0 && axis != 1) || span < 1) return M2D_ERR_ARG;
    if (h->tiles_in_use != 0) { h->err = "m2d_set_shard: the map already holds tiles"; return M2D_ERR_STATE; }
    h->cfg.shard_rank = rank; h->cfg.shard_count = count; h->cfg.shard_axis = axis; h->cfg.shard_span = span;
    h->shard_origin = origin;
    return M2D_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Ingest seam — SURVEY.md §8(f) N4.  In the reference the tracker thread converts the frame BGRA -> BGR
// (GSLAM-DIYSLAM/src/zhaoyong/TrackerOpt.cpp:374-383) and hands (image, pose) to Map2DFusion through a bounded queue
// that drops the OLDEST entry when full (src/DataTrans.h:54-68, capacity 30); Map2DCPU's own worker queue behaves the
// same with capacity 20 (Map2DCPU.cpp:139-142).  Here: m2d_ingest_push() copies/converts into a pinned slot and never
// blocks; one worker thread pops up to kIngestBatch frames at a time and feeds them in order (grouped launches).
// ---------------------------------------------------------------------------------------------------------
static void ingest_worker(m2d_map* m) {
    Ingest& I = *m->ingest;
    std::vector<IngestItem> batch;
    std::vector<const uint8_t*> ptrs;
    std::vector<double> poses;
    std::vector<int> res;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(I.mu);
            I.cv.wait(lk, [&] { return I.stop || (!I.paused && !I.q.empty()); });
            if (I.paused || I.q.empty()) {
                if (I.stop) break;
                continue;
            }
            batch.clear();
            while (!I.q.empty() && (int)batch.size() < kIngestBatch) { batch.push_back(I.q.front()); I.q.pop_front(); }
            I.busy = true;
        }
        ptrs.resize(batch.size()); poses.resize(batch.size() * 7); res.assign(batch.size(), M2D_REJECTED);
        for (size_t i = 0; i < batch.size(); i++) {
            ptrs[i] = I.ring + (size_t)batch[i].slot * I.slot_bytes;
            memcpy(&poses[7 * i], batch[i].pose, sizeof(double) * 7);
        }
        int rc, fused = 0;
        {
            std::lock_guard<std::recursive_mutex> g(m->api);
            rc = m->feed_frames((int)batch.size(), nullptr, 0, I.w, I.h, (size_t)I.w * 3, poses.data(), false, res.data(), ptrs.data());
            int rs = m->sync();   // the slots are read by DMA until here
            if (rc >= 0 && rs < 0) rc = rs;
        }
        for (int r : res) fused += r == M2D_OK;
        {
            std::lock_guard<std::mutex> lk(I.mu);
            for (const IngestItem& b : batch) I.free_slots.push_back(b.slot);
            I.fed += batch.size();
            I.fused += (uint64_t)fused;
            if (rc < 0) I.last_rc = rc;
            I.busy = false;
        }
        I.cv_idle.notify_all();
    }
}

int m2d_ingest_open(m2d_handle h, int capacity, int start_paused) { return m2d_ingest_open_seeded(h, capacity, 0, start_paused); }

int m2d_ingest_open_seeded(m2d_handle h, int capacity, int seed_frames, int start_paused) {
    if (!h || capacity < 1 || capacity > 4096 || seed_frames < 0 || seed_frames > 4096) return M2D_ERR_ARG;
    if (h->ingest) return M2D_ERR_STATE;
    if (h->type == M2D_TYPE_RENDER) { h->err = "m2d_ingest_open: TypeRender renders ONE batch (m2d_render_frames), it has no streaming queue"; return M2D_ERR_UNSUPPORTED; }
    if (!h->valid) { h->err = "m2d_ingest_open: prepare() first (the frame size comes from the camera)"; return M2D_ERR_STATE; }
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) return M2D_ERR_CUDA;
    Ingest* I = new Ingest();
    I->capacity = capacity; I->w = (int)h->g.cam_w; I->h = (int)h->g.cam_h;
    I->slot_bytes = ((size_t)I->w * I->h * 3 + 255) & ~(size_t)255;
    // the queue may hold max(capacity, seed_frames) + 1 frames for an instant (push, then drop one: Map2DCPU.cpp:141-142)
    I->n_slots = std::max(capacity, seed_frames) + 1 + kIngestBatch + kIngestSpare;
    if (cudaHostAlloc((void**)&I->ring, I->slot_bytes * I->n_slots, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        delete I;
        h->err = "m2d_ingest_open: pinned ring allocation failed";
        return M2D_ERR_NOMEM;
    }
    for (int i = I->n_slots - 1; i >= 0; i--) I->free_slots.push_back(i);
    I->paused = start_paused != 0;
    h->ingest = I;
    I->worker = std::thread(ingest_worker, h);
    return M2D_OK;
}

int m2d_ingest_push(m2d_handle h, const uint8_t* pixels, int w, int hpx, size_t stride, int channels, const double* pose) {
    if (!h || !pixels || !pose || (channels != 3 && channels != 4)) return M2D_ERR_ARG;
    Ingest* I = h->ingest;
    if (!I) return M2D_ERR_STATE;
    if (w != I->w || hpx != I->h) {                       // Map2DCPU.cpp:158-162, rejected at the seam already
        fprintf(stderr, "Map2DB200::ingest: frame size != camera size\n");
        return M2D_REJECTED;
    }
    if (stride < (size_t)w * channels) return M2D_ERR_ARG;
    int slot = -1;
    {
        std::lock_guard<std::mutex> lk(I->mu);
        if (I->stop) return M2D_ERR_STATE;
        if (I->free_slots.empty() && !I->q.empty()) {     // ring exhausted (many producers mid-copy): make room like a full queue does
            I->free_slots.push_back(I->q.front().slot);
            I->q.pop_front();
            I->dropped++;
        }
        if (I->free_slots.empty()) return M2D_ERR_STATE;  // more than kIngestSpare producers mid-copy
        slot = I->free_slots.back();
        I->free_slots.pop_back();
    }
    uint8_t* dst = I->ring + (size_t)slot * I->slot_bytes;
    for (int y = 0; y < hpx; y++) {
        const uint8_t* s = pixels + (size_t)y * stride;
        uint8_t* d = dst + (size_t)y * w * 3;
        if (channels == 3) memcpy(d, s, (size_t)w * 3);
        else for (int x = 0; x < w; x++) { d[3 * x] = s[4 * x]; d[3 * x + 1] = s[4 * x + 1]; d[3 * x + 2] = s[4 * x + 2]; }  // CV_BGRA2BGR
    }
    {
        std::lock_guard<std::mutex> lk(I->mu);
        IngestItem it;
        it.slot = slot;
        memcpy(it.pose, pose, sizeof it.pose);
        I->q.push_back(it);
        I->pushed++;
        // Map2DCPU.cpp:141-142 (and DataTrans.h:57-64): push, then drop ONE oldest entry if the queue is over capacity; a queue
        // seeded with more than `capacity` prepare-frames therefore stays that long until the worker catches up
        if ((int)I->q.size() > I->capacity) {
            I->free_slots.push_back(I->q.front().slot);
            I->q.pop_front();
            I->dropped++;
        }
    }
    I->cv.notify_one();
    return M2D_OK;
}

int m2d_ingest_pause(m2d_handle h, int paused) {
    if (!h) return M2D_ERR_ARG;
    Ingest* I = h->ingest;
    if (!I) return M2D_ERR_STATE;
    { std::lock_guard<std::mutex> lk(I->mu); I->paused = paused != 0; }
    I->cv.notify_all();
    return M2D_OK;
}

int m2d_ingest_drain(m2d_handle h) {
    if (!h) return M2D_ERR_ARG;
    Ingest* I = h->ingest;
    if (!I) return M2D_OK;
    std::unique_lock<std::mutex> lk(I->mu);
    if (I->paused) { h->err = "m2d_ingest_drain: the queue is paused"; return M2D_ERR_STATE; }
    I->cv_idle.wait(lk, [&] { return I->q.empty() && !I->busy; });
    return I->last_rc < 0 ? I->last_rc : M2D_OK;
}

static int ingest_shutdown(m2d_handle h, bool discard) {
    if (!h) return M2D_ERR_ARG;
    Ingest* I = h->ingest;
    if (!I) return M2D_OK;
    {
        std::lock_guard<std::mutex> lk(I->mu);
        if (discard) {   // frames still queued are dropped unrendered (what a second prepare() does to the old queue, Map2DCPU.cpp:112-114)
            I->dropped += I->q.size();
            for (const IngestItem& it : I->q) I->free_slots.push_back(it.slot);
            I->q.clear();
        }
        I->paused = false; I->stop = true;   // the worker drains what is queued, then exits
    }
    I->cv.notify_all();
    if (I->worker.joinable()) I->worker.join();
    int rc = I->last_rc;
    {
        std::lock_guard<std::recursive_mutex> g(h->api);
        h->ingest = nullptr;
    }
    cudaSetDevice(h->cfg.device);
    cudaFreeHost(I->ring);
    delete I;
    return rc < 0 ? rc : M2D_OK;
}
int m2d_ingest_close(m2d_handle h) { return ingest_shutdown(h, false); }
int m2d_ingest_abort(m2d_handle h) { return ingest_shutdown(h, true); }

int m2d_ingest_stats(m2d_handle h, uint64_t* pushed, uint64_t* dropped, uint64_t* fed, uint64_t* fused) {
    if (!h) return M2D_ERR_ARG;
    Ingest* I = h->ingest;
    if (!I) return M2D_ERR_STATE;
    std::lock_guard<std::mutex> lk(I->mu);
    if (pushed) *pushed = I->pushed;
    if (dropped) *dropped = I->dropped;
    if (fed) *fed = I->fed;
    if (fused) *fused = I->fused;
    return M2D_OK;
}

int m2d_sync(m2d_handle h) { API_LOCK(h); return h ? h->sync() : M2D_ERR_ARG; }
int m2d_queue_size(m2d_handle h) {
    if (!h) return 0;
    int q = 0;
    if (Ingest* I = h->ingest) {   // frames still queued or in the worker's hands count as "not yet rendered" (Map2D.h:97)
        std::lock_guard<std::mutex> lk(I->mu);
        q = (int)I->q.size();
        if (I->busy) return q + 1;   // the worker holds the handle: do not wait for it, report its batch as one pending unit
    }
    API_LOCK(h);
    return q + h->queue_size();
}
int m2d_set_stream(m2d_handle h, void* s) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h) return M2D_ERR_ARG;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    if (s) {
        if (h->own_stream) cudaStreamDestroy(h->stream);
        h->stream = (cudaStream_t)s;
        h->own_stream = false;
    } else if (!h->own_stream) {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return M2D_ERR_CUDA;
        h->own_stream = true;
    }
    return M2D_OK;
}
int m2d_reset(m2d_handle h) { API_LOCK(h); return h ? h->reset() : M2D_ERR_ARG; }
int m2d_set_input_event(m2d_handle h, void* cuda_event) {
    API_LOCK(h);
    if (h && !h->subs.empty()) { for (m2d_map* sub : h->subs) sub->input_event = (cudaEvent_t)cuda_event; return M2D_OK; }
    if (!h) return M2D_ERR_ARG;
    h->input_event = (cudaEvent_t)cuda_event;
    return M2D_OK;
}

int m2d_get_grid(m2d_handle h, int* w, int* ht, double* mn, double* mx, double* lp) {
    API_LOCK(h);
    if (!h) return M2D_ERR_ARG;
    if (!h->valid) return M2D_ERR_STATE;
    if (w) *w = h->g.w;
    if (ht) *ht = h->g.h;
    if (mn) { mn[0] = h->g.min_x; mn[1] = h->g.min_y; mn[2] = h->min_z; }
    if (mx) { mx[0] = h->g.max_x; mx[1] = h->g.max_y; mx[2] = h->max_z; }
    if (lp) *lp = h->length_pixel;
    return M2D_OK;
}
int m2d_last_rect(m2d_handle h, int* rect) {
    API_LOCK(h);
    if (!h || !rect) return M2D_ERR_ARG;
    memcpy(rect, h->last_rect, sizeof(int) * 4);
    return M2D_OK;
}

int m2d_get_tile(m2d_handle h, int tx, int ty, int level, void* data, float* weight) {
    API_LOCK(h);
    if (h && !h->subs.empty()) {   // the tile lives on the device that owns it
        for (m2d_map* sub : h->subs) { int rc = m2d_get_tile(sub, tx, ty, level, data, weight); if (rc != M2D_REJECTED) { if (rc < 0) h->err = sub->err; return rc; } }
        return M2D_REJECTED;
    }
    if (!h || !data) return M2D_ERR_ARG;
    m2d_map& m = *h;
    std::string& err = m.err;
    if (!m.valid || tx < 0 || ty < 0 || tx >= m.g.w || ty >= m.g.h) return M2D_ERR_ARG;
    const uint8_t* t = m.table[(size_t)ty * m.g.w + tx];
    if (!t) return M2D_REJECTED;
    CU(cudaSetDevice(m.cfg.device));
    if (m.type != M2D_TYPE_MULTIBAND) {
        if (level != 0) return M2D_ERR_ARG;
        CU(cudaMemcpyAsync(data, t, (size_t)kEle * kEle * 4, cudaMemcpyDeviceToHost, m.stream));
        CU(cudaStreamSynchronize(m.stream));
        return M2D_OK;
    }
    if (level < 0 || level >= m.levels) return M2D_ERR_ARG;
    size_t n = (size_t)(kEle >> level), px = n * n;
    std::vector<int16_t> planes(px * 3);
    CU(cudaMemcpyAsync(planes.data(), t + m.lay.lap_off[level], px * 3 * sizeof(int16_t), cudaMemcpyDeviceToHost, m.stream));
    if (weight) CU(cudaMemcpyAsync(weight, t + m.lay.wgt_off[level], px * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
    CU(cudaStreamSynchronize(m.stream));
    int16_t* o = (int16_t*)data;  // planar state -> the reference's interleaved CV_16SC3
    for (size_t i = 0; i < px; i++) { o[3 * i] = planes[i]; o[3 * i + 1] = planes[px + i]; o[3 * i + 2] = planes[2 * px + i]; }
    return M2D_OK;
}

int m2d_get_image(m2d_handle h, uint8_t* out, int* w, int* hpx, int* channels, int* tmx, int* tmy) {
    API_LOCK(h);
    if (!h || !w || !hpx || !channels || !tmx || !tmy) return M2D_ERR_ARG;
    if (out && h->ingest) {
        // The ingest worker may have grown the mosaic since the caller's size query: *w x *h x *channels (as returned by
        // that query) is then the capacity of `out`.  Refuse instead of overrunning it; the caller queries again.
        long long cap = (long long)*w * *hpx * *channels;
        int qw, qh, qc, qx, qy;
        int rc = h->get_image(nullptr, &qw, &qh, &qc, &qx, &qy);
        if (rc != M2D_OK) return rc;
        if ((long long)qw * qh * qc > cap) { h->err = "m2d_get_image: the mosaic grew since the size query; query again"; return M2D_ERR_STATE; }
    }
    return h->get_image(out, w, hpx, channels, tmx, tmy);
}

// ---------------------------------------------------------------------------------------------------------
// Map2DRender (TypeRender = 4) -- Map2DRender.cpp:479-760 renderFrames + the inline MultiBandBlender (:52-310), without the
// GUI (imshow / waitKey) and the seam finder (Map2DRender.EnableSeam = 0).  Host part: the per-frame geometry in FP64 (the
// same operations in the same order as the reference, so that sizes, corners and homographies are its bits), spreadMap, the
// blender's sub-image bookkeeping (:103-139); device part: kernels_render.cu.
// ---------------------------------------------------------------------------------------------------------
int m2d_map::render_frames(int n, const uint8_t* base, size_t frame_stride, int w, int h, size_t stride, const double* poses, bool on_device,
                           int* result) {
    if (type != M2D_TYPE_RENDER || !subs.empty()) return M2D_ERR_STATE;
    for (int i = 0; i < n && result; i++) result[i] = M2D_REJECTED;
    stats.frames_fed += n;
    if (!valid) return M2D_REJECTED;
    if (!base || !poses || n <= 0) return M2D_ERR_ARG;
    if (w != g.cam_w || h != g.cam_h) return M2D_REJECTED;
    if (stride < (size_t)w * 3) return M2D_ERR_ARG;
    CU(cudaSetDevice(cfg.device));
    rnd_have = false;
    if (rnd_wimg_w != w || rnd_wimg_h != h) {
        CU(cudaStreamSynchronize(stream));
        if (rnd_wimg) { CU(cudaFree(rnd_wimg)); rnd_wimg = nullptr; }
        CU(cudaMalloc(&rnd_wimg, (size_t)w * h + 16));
        LAUNCHK(M2D_K_RENDER, launch_rnd_weight_image(w, h, rnd_wimg, stream));
        rnd_wimg_w = w; rnd_wimg_h = h;
    }
    // 1. per frame: ground quad, own bounding box, size, homography into the box (:536-603)
    struct Geo { bool ok; int iw, ih; float cwx, cwy; double hinv[9]; };
    std::vector<Geo> geo((size_t)n);
    double gminx = 0, gminy = 0, gmaxx = 0, gmaxy = 0;   // pi::Point2d min,max start at (0,0) (:532)
    const double lpi = g.length_pixel_inv;
    for (int idx = 0; idx < n; idx++) {
        Geo& F = geo[idx];
        F.ok = false;
        Pose f = pose_mul(g.plane_inv, pose_from7(poses + 7 * (size_t)idx));
        const double ipx[4] = {0, g.cam_w, 0, g.cam_w}, ipy[4] = {0, 0, g.cam_h, g.cam_h};
        double px[4], py[4];
        const double down_z = (f.t.z < 0) ? 1.0 : -1.0;
        bool ok = true;
        for (int j = 0; j < 4; j++) {
            Vec3 u; u.x = (ipx[j] - g.cx) * g.fxinv; u.y = (ipy[j] - g.cy) * g.fyinv; u.z = 1.;
            Vec3 axis = qrot(f.r, u);
            if (axis.x * 0.0 + axis.y * 0.0 + axis.z * down_z < 0.4) { ok = false; break; }
            double s = f.t.z / axis.z;
            px[j] = f.t.x - axis.x * s; py[j] = f.t.y - axis.y * s;
        }
        if (!ok) continue;
        double cminx = 1e6, cminy = 1e6, cmaxx = -1e6, cmaxy = -1e6;
        for (int i = 0; i < 4; i++) {
            if (px[i] < cminx) cminx = px[i];
            if (py[i] < cminy) cminy = py[i];
            if (px[i] > cmaxx) cmaxx = px[i];
            if (py[i] > cmaxy) cmaxy = py[i];
        }
        if (cminx < gminx) gminx = cminx;
        if (cminy < gminy) gminy = cminy;
        if (cmaxx > gmaxx) gmaxx = cmaxx;
        if (cmaxy > gmaxy) gmaxy = cmaxy;
        F.cwx = (float)cminx; F.cwy = (float)cminy;
        F.iw = (int)((cmaxx - cminx) * lpi);
        F.ih = (int)((cmaxy - cminy) * lpi);
        if (F.iw <= 0 || F.ih <= 0) continue;
        // cv::warpPerspective walks its destination in 64-px column blocks once it is >= 64 x 16 px (BLOCK_SZ 32: bh0 = min(16, h),
        // bw0 = min(1024 / bh0, w)); the kernels assume that block base.  A warped frame smaller than that is not a survey frame.
        if (F.iw < 64 || F.ih < 16) { err = "m2d_render_frames: a frame's footprint is below 64 x 16 px at this map scale"; if (result) result[idx] = M2D_ERR_UNSUPPORTED; continue; }
        float srcp[8], dstp[8];
        for (int i = 0; i < 4; i++) {
            srcp[2 * i] = (float)ipx[i]; srcp[2 * i + 1] = (float)ipy[i];
            dstp[2 * i] = (float)((px[i] - cminx) * lpi);
            dstp[2 * i + 1] = (float)((py[i] - cminy) * lpi);
        }
        double M[9];
        if (!perspective_from_points(srcp, dstp, M) || !invert3x3(M, F.hinv)) continue;
        F.ok = true;
        if (result) result[idx] = M2D_OK;
    }
    // 2. spreadMap + whole tiles (:606-639)
    if (gminx < g.min_x || gminy < g.min_y || gmaxx > g.max_x || gmaxy > g.max_y) {
        int rc = spread(gminx, gminy, gmaxx, gmaxy);
        if (rc != M2D_OK) return rc;
    }
    const int xminInt = (int)floor((gminx - g.min_x) * g.ele_size_inv), yminInt = (int)floor((gminy - g.min_y) * g.ele_size_inv);
    const int xmaxInt = (int)ceil((gmaxx - g.min_x) * g.ele_size_inv), ymaxInt = (int)ceil((gmaxy - g.min_y) * g.ele_size_inv);
    if (xminInt < 0 || yminInt < 0 || xmaxInt > g.w || ymaxInt > g.h || xminInt >= xmaxInt || yminInt >= ymaxInt) return M2D_REJECTED;
    const double minx = g.min_x + g.ele_size * xminInt, miny = g.min_y + g.ele_size * yminInt;
    if ((long long)(xmaxInt - xminInt) * (ymaxInt - yminInt) > (1ll << 14)) { err = "m2d_render_frames: canvas larger than 16384 tiles"; return M2D_ERR_UNSUPPORTED; }
    const int Wf = (xmaxInt - xminInt) * kEle, Hf = (ymaxInt - yminInt) * kEle;
    // 3. the blender: band count (:707-715), prepare (:73-101)
    int nb;
    {
        double blend_strength = 5;
        float blend_width = std::sqrt(static_cast<float>(Wf * Hf)) * blend_strength / 100.f;
        nb = static_cast<int>(std::ceil(std::log(blend_width) / std::log(2.)) - 1.);
        if (cfg.render_bands > 0) nb = cfg.render_bands;
        double max_len = (double)std::max(Wf, Hf);
        nb = std::min(nb, (int)std::ceil(std::log(max_len) / std::log(2.0)));
    }
    if (nb < 1 || nb + 1 > kRenderMaxLevels) { err = "m2d_render_frames: band count outside [1, 15]"; return M2D_ERR_UNSUPPORTED; }
    const int step = 1 << nb;
    const int W = Wf + (step - Wf % step) % step, H = Hf + (step - Hf % step) % step;
    const int s16w = cfg.render_blend == 2;
    const size_t wbytes = s16w ? 2 : 4;
    RenderCanvas C{};
    C.levels = nb + 1;
    size_t off = 0, lap_off[kRenderMaxLevels][3], wgt_off[kRenderMaxLevels];
    for (int l = 0, cw = W, ch = H; l <= nb; l++) {
        if (l) { cw = (cw + 1) / 2; ch = (ch + 1) / 2; }
        C.w[l] = cw; C.h[l] = ch;
        const size_t px = (size_t)cw * ch;
        for (int c = 0; c < 3; c++) { lap_off[l][c] = off; off += (px * 2 + 255) & ~(size_t)255; }
        wgt_off[l] = off; off += (px * wbytes + 255) & ~(size_t)255;
    }
    const size_t zero_bytes = off;
    rnd_off16 = off; off += ((size_t)Wf * Hf * 6 + 255) & ~(size_t)255;
    rnd_off8 = off; off += ((size_t)Wf * Hf * 3 + 255) & ~(size_t)255;
    rnd_offmask = off; off += ((size_t)Wf * Hf + 255) & ~(size_t)255;
    CU(cudaStreamSynchronize(stream));
    { int rc = grow((void**)&rnd_canvas, &rnd_canvas_cap, off, false); if (rc != M2D_OK) return rc; }
    for (int l = 0; l <= nb; l++) {
        for (int c = 0; c < 3; c++) C.lap[l][c] = reinterpret_cast<int16_t*>(rnd_canvas + lap_off[l][c]);
        C.wgt[l] = rnd_canvas + wgt_off[l];
    }
    CU(cudaMemsetAsync(rnd_canvas, 0, zero_bytes, stream));
    // the sub-image of every frame (:103-139) and its scratch pyramid
    std::vector<RenderJob> jobs;
    std::vector<int> job_frame;
    std::vector<size_t> job_bytes;
    for (int idx = 0; idx < n; idx++) {
        const Geo& F = geo[idx];
        if (!F.ok) continue;
        const int tlx = (int)((F.cwx - minx) * lpi), tly = (int)((F.cwy - miny) * lpi);   // cornersImages (:648-649)
        const int gap = 3 * step;
        int tnx = std::max(0, tlx - gap), tny = std::max(0, tly - gap);
        int bnx = std::min(W, tlx + F.iw + gap), bny = std::min(H, tly + F.ih + gap);
        tnx = (tnx >> nb) << nb; tny = (tny >> nb) << nb;
        int width = bnx - tnx, height = bny - tny;
        width += (step - width % step) % step;
        height += (step - height % step) % step;
        bnx = tnx + width; bny = tny + height;
        const int dy = std::max(bny - H, 0), dx = std::max(bnx - W, 0);
        tnx -= dx; bnx -= dx; tny -= dy; bny -= dy;
        if (tlx < 0 || tly < 0 || tnx < 0 || tny < 0 || tlx < tnx || tly < tny) { err = "m2d_render_frames: frame corner outside the canvas"; return M2D_ERR_STATE; }
        RenderJob J{};
        memcpy(J.hinv, F.hinv, sizeof J.hinv);
        J.iw = F.iw; J.ih = F.ih; J.left = tlx - tnx; J.top = tly - tny; J.x_tl = tnx; J.y_tl = tny; J.sw = width; J.sh = height;
        size_t b = 0;
        for (int l = 0; l <= nb; l++) {
            const size_t px = (size_t)(width >> l) * (height >> l);
            J.g_off[l] = b; b += (px * 4 + 255) & ~(size_t)255;
            J.w_off[l] = b; b += (px * wbytes + 255) & ~(size_t)255;
        }
        jobs.push_back(J); job_frame.push_back(idx); job_bytes.push_back(b);
    }
    // 4. chunks of frames that fit the scratch budget, each blended into the canvas in feed order
    const size_t frame_bytes = (size_t)h * stride;
    size_t j0 = 0;
    while (j0 < jobs.size()) {
        size_t j1 = j0, bytes = 0;
        int max_px = 0;
        while (j1 < jobs.size() && j1 - j0 < 1024 && (j1 == j0 || bytes + job_bytes[j1] <= (size_t)scratch_budget)) {
            for (int l = 0; l <= nb; l++) { jobs[j1].g_off[l] += bytes; jobs[j1].w_off[l] += bytes; }
            bytes += job_bytes[j1];
            max_px = std::max(max_px, jobs[j1].sw * jobs[j1].sh);
            j1++;
        }
        const int m = (int)(j1 - j0);
        CU(cudaStreamSynchronize(stream));   // the previous chunk still reads the buffers that may move below
        { int rc = grow((void**)&rnd_scratch, &rnd_scratch_cap, bytes, false); if (rc != M2D_OK) return rc; }
        { int rc = grow((void**)&rnd_jobs, &rnd_jobs_cap, (size_t)m * sizeof(RenderJob), false); if (rc != M2D_OK) return rc; }
        if (!on_device) { int rc = grow((void**)&rnd_raw, &rnd_raw_cap, (size_t)m * frame_bytes + 16, false); if (rc != M2D_OK) return rc; }
        for (int k = 0; k < m; k++) {
            RenderJob& J = jobs[j0 + k];
            const uint8_t* src = base + (size_t)job_frame[j0 + k] * frame_stride;
            J.raw_stride = (int)stride;
            if (on_device) J.raw = src;
            else {
                J.raw = rnd_raw + (size_t)k * frame_bytes;
                CU(cudaMemcpyAsync(rnd_raw + (size_t)k * frame_bytes, src, frame_bytes, cudaMemcpyHostToDevice, stream));
            }
        }
        CU(cudaMemcpyAsync(rnd_jobs, jobs.data() + j0, (size_t)m * sizeof(RenderJob), cudaMemcpyHostToDevice, stream));
        const RenderJob* dj = reinterpret_cast<const RenderJob*>(rnd_jobs);
        LAUNCHK(M2D_K_RENDER, launch_rnd_warp(dj, m, max_px, rnd_scratch, rnd_wimg, w, h, s16w, stream));
        for (int l = 0; l < nb; l++) LAUNCHK(M2D_K_RENDER, launch_rnd_pyrdown(dj, m, max_px, rnd_scratch, l, cfg.f32_mode, s16w, stream));
        LAUNCHK(M2D_K_RENDER, launch_rnd_blend(dj, m, rnd_scratch, C, cfg.render_blend, stream));
        stats.frames_fused += (uint64_t)m;
        stats.input_px += (uint64_t)m * (uint64_t)w * h;
        j0 = j1;
    }
    // 5. blend(): normalise (weighted sums), restore, mask, crop (:256-300)
    if (cfg.render_blend != 0) LAUNCHK(M2D_K_RENDER, launch_rnd_normalize(C, cfg.render_blend, stream));
    for (int l = nb; l > 0; l--) {
        MosaicLevel cl{{C.lap[l][0], C.lap[l][1], C.lap[l][2]}, C.w[l], C.h[l]}, fl{{C.lap[l - 1][0], C.lap[l - 1][1], C.lap[l - 1][2]}, C.w[l - 1], C.h[l - 1]};
        LAUNCHK(M2D_K_COLLAPSE, launch_mosaic_upadd(cl, fl, stream));
    }
    LAUNCHK(M2D_K_RENDER, launch_rnd_final(C, s16w, Wf, Hf, reinterpret_cast<int16_t*>(rnd_canvas + rnd_off16), rnd_canvas + rnd_off8,
                                            rnd_canvas + rnd_offmask, stream));
    rnd_w = Wf; rnd_h = Hf; rnd_bands = nb; rnd_tx0 = xminInt + org_x; rnd_ty0 = yminInt + org_y;
    rnd_have = true;
    last_rect[0] = xminInt; last_rect[1] = yminInt; last_rect[2] = xmaxInt; last_rect[3] = ymaxInt;
    if (!on_device) CU(cudaStreamSynchronize(stream));   // the caller's (pageable or pinned) frames are free again
    return M2D_OK;
}

int m2d_render_frames(m2d_handle h, int n, const uint8_t* base, size_t frame_stride, int w, int hpx, size_t stride, const double* poses,
                      int on_device, int* result) {
    API_LOCK(h);
    if (!h) return M2D_ERR_ARG;
    return h->render_frames(n, base, frame_stride, w, hpx, stride, poses, on_device != 0, result);
}

int m2d_render_get(m2d_handle h, int16_t* result16, uint8_t* mask, int* w, int* hpx, int* num_bands, int* tile_x0, int* tile_y0) {
    API_LOCK(h);
    if (!h || h->type != M2D_TYPE_RENDER) return M2D_ERR_ARG;
    m2d_map& m = *h;
    std::string& err = m.err;
    if (!m.rnd_have) return M2D_REJECTED;
    if (w) *w = m.rnd_w;
    if (hpx) *hpx = m.rnd_h;
    if (num_bands) *num_bands = m.rnd_bands;
    if (tile_x0) *tile_x0 = m.rnd_tx0;
    if (tile_y0) *tile_y0 = m.rnd_ty0;
    if (!result16 && !mask) return M2D_OK;
    CU(cudaSetDevice(m.cfg.device));
    const size_t px = (size_t)m.rnd_w * m.rnd_h;
    if (result16) CU(cudaMemcpyAsync(result16, m.rnd_canvas + m.rnd_off16, px * 6, cudaMemcpyDeviceToHost, m.stream));
    if (mask) CU(cudaMemcpyAsync(mask, m.rnd_canvas + m.rnd_offmask, px, cudaMemcpyDeviceToHost, m.stream));
    CU(cudaStreamSynchronize(m.stream));
    return M2D_OK;
}

int m2d_tile_bbox(m2d_handle h, int* bbox_abs) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || !bbox_abs) return M2D_ERR_ARG;
    if (!h->valid) return M2D_ERR_STATE;
    int x0, y0, x1, y1;
    if (!h->tile_bbox(x0, y0, x1, y1)) return M2D_REJECTED;
    bbox_abs[0] = x0 + h->org_x; bbox_abs[1] = y0 + h->org_y; bbox_abs[2] = x1 + h->org_x; bbox_abs[3] = y1 + h->org_y;
    return M2D_OK;
}

int m2d_get_image_rect(m2d_handle h, uint8_t* out, int out_on_device, const int* window_abs, const int* crop_abs, int* w, int* hpx,
                       int* channels) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || !window_abs || !crop_abs || !w || !hpx || !channels) return M2D_ERR_ARG;
    if (!h->valid) return M2D_ERR_STATE;
    const int ox = h->org_x, oy = h->org_y;
    const int win[4] = {window_abs[0] - ox, window_abs[1] - oy, window_abs[2] - ox, window_abs[3] - oy};
    const int crop[4] = {crop_abs[0] - ox, crop_abs[1] - oy, crop_abs[2] - ox, crop_abs[3] - oy};
    return h->collapse_window(out, out_on_device != 0, win, crop, w, hpx, channels);
}

int m2d_drop_tiles_rect(m2d_handle h, const int* rect_abs, int* n_dropped) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || !rect_abs) return M2D_ERR_ARG;
    m2d_map& m = *h;
    if (!m.valid) return M2D_ERR_STATE;
    if (cudaSetDevice(m.cfg.device) != cudaSuccess || cudaStreamSynchronize(m.stream) != cudaSuccess) return M2D_ERR_CUDA;
    int n = 0;
    std::lock_guard<std::mutex> lk(m.pool_mu);
    for (int y = std::max(rect_abs[1] - m.org_y, 0); y < std::min(rect_abs[3] - m.org_y, m.g.h); y++)
        for (int x = std::max(rect_abs[0] - m.org_x, 0); x < std::min(rect_abs[2] - m.org_x, m.g.w); x++) {
            uint8_t*& t = m.table[(size_t)y * m.g.w + x];
            if (t) { m.free_tiles.push_back(t); t = nullptr; m.tiles_in_use--; n++; }
        }
    if (n_dropped) *n_dropped = n;
    return M2D_OK;
}

int m2d_save(m2d_handle h, const char* filename) {
    API_LOCK(h);
    if (!h || !filename) return M2D_ERR_ARG;
    int w, hp, cn, tx, ty;
    int rc = h->get_image(nullptr, &w, &hp, &cn, &tx, &ty);
    if (rc != M2D_OK) return rc;
    std::vector<uint8_t> img((size_t)w * hp * cn);
    rc = h->get_image(img.data(), &w, &hp, &cn, &tx, &ty);
    if (rc != M2D_OK) return rc;
    if (m2d_write_png(filename, img.data(), w, hp, cn) != 0) {
        h->err = std::string("cannot write ") + filename;
        return M2D_ERR_IO;
    }
    return M2D_OK;
}

size_t m2d_tile_bytes(m2d_handle h) { return h ? h->tile_bytes : 0; }
size_t m2d_tile_state_bytes(m2d_handle h) {
    if (!h) return 0;
    return h->type == M2D_TYPE_MULTIBAND ? h->lay.cmin_off : h->tile_bytes;
}
int m2d_tile_count(m2d_handle h) {
    API_LOCK(h);
    if (!h) return 0;
    if (!h->subs.empty()) { int n = 0; for (m2d_map* sub : h->subs) n += (int)sub->tiles_in_use; return n; }
    return (int)h->tiles_in_use;
}

int m2d_export_tiles(m2d_handle h, int max_tiles, int* abs_xy, uint8_t* dst, int dst_on_device, int* n_out) {
    return m2d_export_tiles_rect(h, nullptr, max_tiles, abs_xy, dst, dst_on_device, n_out);
}

int m2d_export_tiles_rect(m2d_handle h, const int* rect_abs, int max_tiles, int* abs_xy, uint8_t* dst, int dst_on_device, int* n_out) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || !abs_xy || !n_out || (!dst && max_tiles > 0)) return M2D_ERR_ARG;
    m2d_map& m = *h;
    std::string& err = m.err;
    uint64_t& launches = m.launches;
    bool& profiling = m.profiling;
    std::vector<ProfRec>& prof = m.prof;
    cudaStream_t stream = m.stream;
    if (!m.valid) return M2D_ERR_STATE;
    CU(cudaSetDevice(m.cfg.device));
    std::vector<uint8_t*> ptrs;
    int ry0 = 0, ry1 = m.g.h, rx0 = 0, rx1 = m.g.w;
    if (rect_abs) {
        rx0 = std::max(rect_abs[0] - m.org_x, 0); ry0 = std::max(rect_abs[1] - m.org_y, 0);
        rx1 = std::min(rect_abs[2] - m.org_x, m.g.w); ry1 = std::min(rect_abs[3] - m.org_y, m.g.h);
    }
    for (int y = ry0; y < ry1; y++)
        for (int x = rx0; x < rx1; x++) {
            uint8_t* t = m.table[(size_t)y * m.g.w + x];
            if (!t) continue;
            if ((int)ptrs.size() >= max_tiles) { if (max_tiles == 0) { ptrs.push_back(t); continue; } return M2D_ERR_ARG; }
            abs_xy[2 * ptrs.size()] = x + m.org_x; abs_xy[2 * ptrs.size() + 1] = y + m.org_y;
            ptrs.push_back(t);
        }
    int n = (int)ptrs.size();
    *n_out = n;
    if (n == 0 || max_tiles == 0) return M2D_OK;   // max_tiles == 0: count only
    if (dst_on_device && (m.tile_bytes % 16) == 0 && (reinterpret_cast<uintptr_t>(dst) % 16) == 0) {
        // one gather kernel instead of n small copies
        CU(cudaStreamSynchronize(stream));
        { int rc = m.grow((void**)&m.d_collapse, &m.collapse_cap, (size_t)n * sizeof(uint8_t*) + 256, false); if (rc != M2D_OK) return rc; }
        CU(cudaMemcpyAsync(m.d_collapse, ptrs.data(), (size_t)n * sizeof(uint8_t*), cudaMemcpyHostToDevice, stream));
        LAUNCHK(M2D_K_MISC, launch_tile_copy(reinterpret_cast<uint8_t* const*>(m.d_collapse), n, dst, m.tile_bytes, 1, stream));
    } else {
        for (int i = 0; i < n; i++)
            CU(cudaMemcpyAsync(dst + (size_t)i * m.tile_bytes, ptrs[i], m.tile_bytes,
                               dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    }
    CU(cudaStreamSynchronize(stream));
    return M2D_OK;
}

int m2d_import_tiles(m2d_handle h, int n, const int* abs_xy, const uint8_t* src, int src_on_device) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || n < 0 || (n && (!abs_xy || !src))) return M2D_ERR_ARG;
    m2d_map& m = *h;
    std::string& err = m.err;
    uint64_t& launches = m.launches;
    bool& profiling = m.profiling;
    std::vector<ProfRec>& prof = m.prof;
    cudaStream_t stream = m.stream;
    if (!m.valid) return M2D_ERR_STATE;
    if (n == 0) return M2D_OK;
    CU(cudaSetDevice(m.cfg.device));
    std::vector<uint8_t*> ptrs(n);
    for (int i = 0; i < n; i++) {
        int x = abs_xy[2 * i] - m.org_x, y = abs_xy[2 * i + 1] - m.org_y;
        if (x < 0 || y < 0 || x >= m.g.w || y >= m.g.h) { err = "import: tile outside the grid (shards must see the same poses)"; return M2D_ERR_ARG; }
        uint8_t*& slot = m.table[(size_t)y * m.g.w + x];
        if (!slot) { int rc = m.alloc_tile(&slot); if (rc != M2D_OK) return rc; }
        ptrs[i] = slot;
    }
    if (src_on_device && (m.tile_bytes % 16) == 0 && (reinterpret_cast<uintptr_t>(src) % 16) == 0) {
        CU(cudaStreamSynchronize(stream));
        { int rc = m.grow((void**)&m.d_collapse, &m.collapse_cap, (size_t)n * sizeof(uint8_t*) + 256, false); if (rc != M2D_OK) return rc; }
        CU(cudaMemcpyAsync(m.d_collapse, ptrs.data(), (size_t)n * sizeof(uint8_t*), cudaMemcpyHostToDevice, stream));
        LAUNCHK(M2D_K_MISC, launch_tile_copy(reinterpret_cast<uint8_t* const*>(m.d_collapse), n, const_cast<uint8_t*>(src), m.tile_bytes, 0, stream));
    } else {
        for (int i = 0; i < n; i++)
            CU(cudaMemcpyAsync(ptrs[i], src + (size_t)i * m.tile_bytes, m.tile_bytes,
                               src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    }
    CU(cudaStreamSynchronize(stream));
    return M2D_OK;
}

int m2d_poll_changed(m2d_handle h, int max_tiles, int* xy, int* n_out) {
    API_LOCK(h);
    if (h && !h->subs.empty()) {   // every device tracks the tiles it owns
        if (!xy || !n_out || max_tiles < 0) return M2D_ERR_ARG;
        int n = 0;
        for (m2d_map* sub : h->subs) { int k = 0; int rc = m2d_poll_changed(sub, max_tiles - n, xy + 2 * n, &k); if (rc != M2D_OK) return rc; n += k; }
        *n_out = n;
        return M2D_OK;
    }
    if (!h || !xy || !n_out || max_tiles < 0) return M2D_ERR_ARG;
    m2d_map& m = *h;
    if (!m.valid) return M2D_ERR_STATE;
    int n = 0;
    for (int y = 0; y < m.g.h && n < max_tiles; y++)
        for (int x = 0; x < m.g.w && n < max_tiles; x++) {
            size_t gi = (size_t)y * m.g.w + x;
            if (m.changed[gi] && m.table[gi]) { xy[2 * n] = x; xy[2 * n + 1] = y; n++; m.changed[gi] = 0; }
        }
    *n_out = n;
    return M2D_OK;
}

int m2d_get_tile_image(m2d_handle h, int tx, int ty, int high_quality, uint8_t* out, int* channels) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || !out || !channels) return M2D_ERR_ARG;
    m2d_map& m = *h;
    std::string& err = m.err;
    uint64_t& launches = m.launches;
    bool& profiling = m.profiling;
    std::vector<ProfRec>& prof = m.prof;
    cudaStream_t stream = m.stream;
    if (!m.valid || tx < 0 || ty < 0 || tx >= m.g.w || ty >= m.g.h) return M2D_ERR_ARG;
    const uint8_t* t = m.table[(size_t)ty * m.g.w + tx];
    if (!t) return M2D_REJECTED;
    CU(cudaSetDevice(m.cfg.device));
    if (m.type != M2D_TYPE_MULTIBAND) {
        *channels = 4;
        CU(cudaMemcpyAsync(out, t, (size_t)kEle * kEle * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return M2D_OK;
    }
    *channels = 3;
    const int levels = m.levels;
    const uint8_t* nb[9];
    bool all9 = high_quality != 0;
    for (int yi = ty - 1, k = 0; yi <= ty + 1; yi++)
        for (int xi = tx - 1; xi <= tx + 1; xi++, k++) {
            nb[k] = (yi < 0 || yi >= m.g.h || xi < 0 || xi >= m.g.w) ? nullptr : m.table[(size_t)yi * m.g.w + xi];
            if (!nb[k]) all9 = false;
        }
    MosaicSet ms{};
    std::vector<SubPaste> items;
    size_t off = 0, offs[M2D_MAX_LEVELS][3];
    int b0 = 0;
    for (int i = 0; i < levels; i++) {
        int n = kEle >> i, b = all9 ? (1 << (levels - i - 1)) : 0, d = n + 2 * b;
        if (i == 0) b0 = b;
        ms.lv[i].w = ms.lv[i].h = d;
        for (int c = 0; c < 3; c++) { offs[i][c] = off; off += ((size_t)d * d * sizeof(int16_t) + 255) & ~(size_t)255; }
        for (int y = 0; y < 3; y++)
            for (int x = 0; x < 3; x++) {
                if (!all9 && !(x == 1 && y == 1)) continue;
                SubPaste sp;
                sp.tile = all9 ? nb[3 * y + x] : t;
                sp.level = i;
                sp.w = (x == 1) ? n : b; sp.h = (y == 1) ? n : b;
                sp.sx = (x == 0) ? (n - b) : 0; sp.sy = (y == 0) ? (n - b) : 0;
                sp.dx = (x == 0) ? 0 : ((x == 1) ? b : (d - b)); sp.dy = (y == 0) ? 0 : ((y == 1) ? b : (d - b));
                if (sp.w > 0 && sp.h > 0) items.push_back(sp);
            }
    }
    size_t zero_bytes = off;
    size_t off_out = off; off += (size_t)kEle * kEle * 3 + 256;
    size_t off_items = (off + 255) & ~(size_t)255; off = off_items + items.size() * sizeof(SubPaste) + 256;
    CU(cudaStreamSynchronize(stream));
    { int rc = m.grow((void**)&m.d_collapse, &m.collapse_cap, off, false); if (rc != M2D_OK) return rc; }
    for (int i = 0; i < levels; i++)
        for (int c = 0; c < 3; c++) ms.lv[i].g[c] = reinterpret_cast<int16_t*>(m.d_collapse + offs[i][c]);
    ms.w0 = nullptr;
    SubPaste* d_items = reinterpret_cast<SubPaste*>(m.d_collapse + off_items);
    uint8_t* d_out = m.d_collapse + off_out;
    CU(cudaMemsetAsync(m.d_collapse, 0, zero_bytes, stream));
    CU(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(SubPaste), cudaMemcpyHostToDevice, stream));
    LAUNCHK(M2D_K_COLLAPSE, launch_sub_paste(d_items, (int)items.size(), m.lay, ms, stream));
    for (int l = levels - 1; l > 0; l--) LAUNCHK(M2D_K_COLLAPSE, launch_mosaic_upadd(ms.lv[l], ms.lv[l - 1], stream));
    LAUNCHK(M2D_K_COLLAPSE, launch_tile_crop(ms.lv[0], b0, reinterpret_cast<const float*>(t + m.lay.wgt_off[0]), d_out, stream));
    CU(cudaMemcpyAsync(out, d_out, (size_t)kEle * kEle * 3, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return M2D_OK;
}

namespace {
struct StateHeader {
    char magic[8];            // "M2DSTATE"
    uint32_t version, type, levels, reserved;
    uint64_t tile_bytes, n_tiles;
    GridGeom g;
    double min_z, max_z, length_pixel;
    int32_t org_x, org_y;
};
}  // namespace

int m2d_save_state(m2d_handle h, const char* filename) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || !filename) return M2D_ERR_ARG;
    m2d_map& m = *h;
    if (!m.valid) return M2D_ERR_STATE;
    int n = (int)m.tiles_in_use;
    std::vector<int> xy((size_t)std::max(n, 1) * 2);
    std::vector<uint8_t> buf((size_t)std::max(n, 1) * m.tile_bytes);
    int got = 0;
    int rc = m2d_export_tiles(h, n, xy.data(), buf.data(), 0, &got);
    if (rc != M2D_OK) return rc;
    StateHeader hd{};
    memcpy(hd.magic, "M2DSTATE", 8);
    hd.version = 1; hd.type = (uint32_t)m.type; hd.levels = (uint32_t)m.levels;
    hd.tile_bytes = m.tile_bytes; hd.n_tiles = (uint64_t)got;
    hd.g = m.g; hd.min_z = m.min_z; hd.max_z = m.max_z; hd.length_pixel = m.length_pixel;
    hd.org_x = m.org_x; hd.org_y = m.org_y;
    FILE* f = fopen(filename, "wb");
    if (!f) { m.err = std::string("cannot open ") + filename; return M2D_ERR_IO; }
    bool ok = fwrite(&hd, sizeof hd, 1, f) == 1 && (got == 0 || fwrite(xy.data(), sizeof(int) * 2, got, f) == (size_t)got) &&
              (got == 0 || fwrite(buf.data(), m.tile_bytes, got, f) == (size_t)got);
    ok = (fclose(f) == 0) && ok;
    if (!ok) { m.err = std::string("short write to ") + filename; return M2D_ERR_IO; }
    return M2D_OK;
}

int m2d_load_state(m2d_handle h, const char* filename) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || !filename) return M2D_ERR_ARG;
    m2d_map& m = *h;
    std::string& err = m.err;
    FILE* f = fopen(filename, "rb");
    if (!f) { err = std::string("cannot open ") + filename; return M2D_ERR_IO; }
    StateHeader hd{};
    if (fread(&hd, sizeof hd, 1, f) != 1 || memcmp(hd.magic, "M2DSTATE", 8) != 0 || hd.version != 1) {
        fclose(f); err = "not a map2d_b200 state file"; return M2D_ERR_IO;
    }
    if ((int)hd.type != m.type || (int)hd.levels != m.levels || hd.tile_bytes != m.tile_bytes) {
        fclose(f); err = "state file was written by a handle of another type / band number"; return M2D_ERR_ARG;
    }
    if (hd.g.w <= 0 || hd.g.h <= 0 || (long long)hd.g.w * hd.g.h > (1ll << 26) || hd.n_tiles > (uint64_t)hd.g.w * hd.g.h) {
        fclose(f); err = "corrupt state header"; return M2D_ERR_IO;
    }
    std::vector<int> xy((size_t)std::max<uint64_t>(hd.n_tiles, 1) * 2);
    std::vector<uint8_t> buf((size_t)std::max<uint64_t>(hd.n_tiles, 1) * m.tile_bytes);
    bool ok = (hd.n_tiles == 0 || fread(xy.data(), sizeof(int) * 2, hd.n_tiles, f) == hd.n_tiles) &&
              (hd.n_tiles == 0 || fread(buf.data(), m.tile_bytes, hd.n_tiles, f) == hd.n_tiles);
    fclose(f);
    if (!ok) { err = "truncated state file"; return M2D_ERR_IO; }
    CU(cudaSetDevice(m.cfg.device));
    if (m.valid) { int rc = m.reset(); if (rc != M2D_OK) return rc; }
    m.g = hd.g; m.min_z = hd.min_z; m.max_z = hd.max_z; m.length_pixel = hd.length_pixel;
    m.org_x = hd.org_x; m.org_y = hd.org_y;
    m.table.assign((size_t)hd.g.w * hd.g.h, nullptr);
    m.slot_work.assign((size_t)hd.g.w * hd.g.h, -1);
    m.slot_epoch.assign((size_t)hd.g.w * hd.g.h, 0);
    m.changed.assign((size_t)hd.g.w * hd.g.h, 1);
    m.last_rect[0] = m.last_rect[1] = m.last_rect[2] = m.last_rect[3] = -1;
    m.valid = true;
    return m2d_import_tiles(h, (int)hd.n_tiles, xy.data(), buf.data(), 0);
}

int m2d_get_stats(m2d_handle h, m2d_stats* out) {
    API_LOCK(h);
    if (h && !h->subs.empty()) return m2d_get_stats(h->subs[0], out);   // frame counters are identical on every device; px counters are the first device's share
    if (!h || !out) return M2D_ERR_ARG;
    m2d_map& m = *h;
    std::string& err = m.err;
    CU(cudaSetDevice(m.cfg.device));
    unsigned long long d[32];
    CU(cudaMemcpyAsync(d, m.d_stats, sizeof d, cudaMemcpyDeviceToHost, m.stream));
    CU(cudaStreamSynchronize(m.stream));
    *out = m.stats;
    if (m.type == M2D_TYPE_MULTIBAND)
        for (int l = 0; l < M2D_MAX_LEVELS; l++) {
            out->win_px[l] = d[l];
            out->need_px[l] = l < 6 ? d[20 + l] : 0;
            out->needw_px[l] = l < 6 ? d[26 + l] : 0;
        }
    else { out->win_px[0] = d[17]; out->footprint_px = d[16]; }
    return M2D_OK;
}

int m2d_profile(m2d_handle h, int enable) {
    API_LOCK(h);
    if (h && !h->subs.empty()) { for (m2d_map* sub : h->subs) sub->profiling = enable != 0; return M2D_OK; }
    if (!h) return M2D_ERR_ARG;
    h->profiling = enable != 0;
    return M2D_OK;
}
int m2d_get_kernel_times(m2d_handle h, double* ms, uint64_t* count) {
    API_LOCK(h);
    MULTI_UNSUPPORTED(h);
    if (!h || !ms || !count) return M2D_ERR_ARG;
    m2d_map& m = *h;
    std::string& err = m.err;
    CU(cudaSetDevice(m.cfg.device));
    CU(cudaStreamSynchronize(m.stream));
    for (int i = 0; i < M2D_KERNEL_CLASSES; i++) { ms[i] = 0; count[i] = 0; }
    for (ProfRec& r : m.prof) {
        float t = 0;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess && r.kind >= 0 && r.kind < M2D_KERNEL_CLASSES) { ms[r.kind] += t; count[r.kind]++; }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    m.prof.clear();
    return M2D_OK;
}

const char* m2d_last_error(m2d_handle h) { return h ? h->err.c_str() : "null handle"; }
uint64_t m2d_launch_count(m2d_handle h) {
    if (!h) return 0;
    uint64_t n = h->launches;
    for (m2d_map* sub : h->subs) n += sub->launches;
    return n;
}

int m2d_tile_gps_corners(const double* plane7, double grid_min_x, double grid_min_y, double ele_size, int tx, int ty,
                         const double* gps_origin, double* tl, double* br) {
    if (!plane7 || !gps_origin || !tl || !br) return M2D_ERR_ARG;
    // MultiBandMap2DCPU.cpp:709-712: the corners live in float variables
    const float x0 = (float)(grid_min_x + tx * ele_size), y0 = (float)(grid_min_y + ty * ele_size);
    const float x1 = (float)(x0 + ele_size), y1 = (float)(y0 + ele_size);
    const Pose plane = pose_from7(plane7);
    const double lng1 = gps_origin[0], lat1 = gps_origin[1];
    // pi::calcLngLatFromDistance, utils_GPS.cpp:133-160 (EARTH_RADIUS 6378137, DEG2RAD 0.017453292519943, f = 1/298.257223563)
    const double a = 6378137.0, deg2rad = 0.017453292519943, f = 1.0 / 298.257223563, e_2 = 2 * f - f * f;
    const double phi = lat1 * deg2rad, sp = sin(phi);
    const double lng_unit = deg2rad * a * cos(phi) / sqrt(1 - e_2 * (sp * sp));
    const double lat_unit = deg2rad * a * (1 - e_2) / pow(1 - e_2 * (sp * sp), 1.5);
    const float cx[2] = {x0, x1}, cy[2] = {y0, y1};
    double* out[2] = {tl, br};
    for (int i = 0; i < 2; i++) {
        Vec3 v{(double)cx[i], (double)cy[i], 0.0};
        Vec3 r = qrot(plane.r, v);                       // SE3 * point = t + R * p  (SE3.h:99-101)
        const double wx = plane.t.x + r.x, wy = plane.t.y + r.y;
        out[i][0] = wx / lng_unit + lng1;
        out[i][1] = wy / lat_unit + lat1;
        out[i][2] = 0.0;
    }
    return M2D_OK;
}

int m2d_cell_weight_bounds(const double* hinv, int nx, int ny, int sw, int sh, int weight_type, int level, int cx, int cy,
                           float* lo, float* hi) {
    if (!hinv || !lo || !hi || level < 0 || level > 5 || nx < 1 || ny < 1 || cx < 0 || cy < 0 || cx >= nx * 8 || cy >= ny * 8) return M2D_ERR_ARG;
    float m[9];
    for (int i = 0; i < 9; i++) m[i] = (float)hinv[i];   // FrameJob::hinvf
    cell_weight_bounds(m, nx, ny, sw, sh, weight_type, level, cx, cy, lo, hi);
    return M2D_OK;
}

int m2d_pull_cell_rect(const double* hinv, int X0, int Y0, int sw, int sh, int* rect) {
    if (!hinv || !rect || sw < 1 || sh < 1) return M2D_ERR_ARG;
    return pull_cell_rect(hinv, X0, Y0, sw, sh, &rect[0], &rect[1], &rect[2], &rect[3]) ? M2D_OK : M2D_REJECTED;
}

int m2d_weight_reach_table(int levels, unsigned char* lo, unsigned char* hi) {
    if (!lo || !hi || levels < 1 || levels > 6) return M2D_ERR_ARG;
    unsigned char l[6][6], h[6][6];
    make_weight_reach_table(levels, l, h);
    memcpy(lo, l, 36);
    memcpy(hi, h, 36);
    return M2D_OK;
}

int m2d_reach_table(int levels, unsigned char* lo, unsigned char* hi) {
    if (!lo || !hi || levels < 1 || levels > 6) return M2D_ERR_ARG;
    unsigned char l[6][6], h[6][6];
    make_reach_table(levels, l, h);
    memcpy(lo, l, 36);
    memcpy(hi, h, 36);
    return M2D_OK;
}

// ---- device buffers that other PROCESSES on the node can map (CUDA IPC): frames stay where they were captured and the
// kernels of a neighbouring rank sample them in place over NVLink -- no halo copies, only the px that are really needed move
void* m2d_device_alloc(int device, size_t bytes) {
    void* p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void m2d_device_free(int device, void* p) {
    if (!p) return;
    cudaSetDevice(device);
    cudaFree(p);
}
int m2d_ipc_export(void* dptr, unsigned char* handle64) {
    if (!dptr || !handle64) return M2D_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, dptr) != cudaSuccess) { cudaGetLastError(); return M2D_ERR_CUDA; }
    memcpy(handle64, &hd, 64);
    return M2D_OK;
}
int m2d_ipc_open(int device, const unsigned char* handle64, void** dptr) {
    if (!handle64 || !dptr) return M2D_ERR_ARG;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    if (cudaSetDevice(device) != cudaSuccess || cudaIpcOpenMemHandle(dptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return M2D_ERR_CUDA; }
    return M2D_OK;
}
int m2d_ipc_close(int device, void* dptr) {
    if (!dptr) return M2D_OK;
    if (cudaSetDevice(device) != cudaSuccess || cudaIpcCloseMemHandle(dptr) != cudaSuccess) { cudaGetLastError(); return M2D_ERR_CUDA; }
    return M2D_OK;
}

void* m2d_alloc_host(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void m2d_free_host(void* p) { if (p) cudaFreeHost(p); }

int m2d_compute_bounds(m2d_handle h, int n, const double* poses, int* rects, double* hinv) {
    API_LOCK(h);
    if (h && !h->subs.empty()) return m2d_compute_bounds(h->subs[0], n, poses, rects, hinv);
    if (!h || n < 0 || !poses || !rects || !hinv) return M2D_ERR_ARG;
    m2d_map& m = *h;
    std::string& err = m.err;
    uint64_t& launches = m.launches;
    bool& profiling = m.profiling;
    std::vector<ProfRec>& prof = m.prof;
    cudaStream_t stream = m.stream;
    if (!m.valid) return M2D_ERR_STATE;
    if (n == 0) return M2D_OK;
    CU(cudaSetDevice(m.cfg.device));
    double* d_poses = nullptr;
    FrameBounds* d_fb = nullptr;
    CU(cudaMalloc(&d_poses, (size_t)n * 7 * sizeof(double)));
    cudaError_t e = cudaMalloc(&d_fb, (size_t)n * sizeof(FrameBounds));
    if (e != cudaSuccess) { cudaFree(d_poses); err = "cudaMalloc"; return M2D_ERR_NOMEM; }
    std::vector<FrameBounds> fb(n);
    auto body = [&]() -> int {
        CU(cudaMemcpyAsync(d_poses, poses, (size_t)n * 7 * sizeof(double), cudaMemcpyHostToDevice, m.stream));
        LAUNCH(launch_bounds(m.g, n, d_poses, d_fb, m.stream));
        CU(cudaMemcpyAsync(fb.data(), d_fb, (size_t)n * sizeof(FrameBounds), cudaMemcpyDeviceToHost, m.stream));
        CU(cudaStreamSynchronize(m.stream));
        return M2D_OK;
    };
    int rc = body();
    cudaFree(d_poses);
    cudaFree(d_fb);
    if (rc != M2D_OK) return rc;
    for (int i = 0; i < n; i++) {
        if (fb[i].ok) { memcpy(rects + 4 * i, &fb[i].x0, 4 * sizeof(int)); memcpy(hinv + 9 * i, fb[i].hinv, 9 * sizeof(double)); }
        else { for (int k = 0; k < 4; k++) rects[4 * i + k] = -1; for (int k = 0; k < 9; k++) hinv[9 * i + k] = 0; }
    }
    return M2D_OK;
}

}  // extern "C"
