/*
 * map2d_b200.h — C-ABI of the B200-native Map2DFusion feed() hot path.
 *
 * This is the drop-in boundary: every entry point below replaces one member of the reference's
 * `Map2D` plugin interface (reference paths are relative to the pi-slam-fusion tree) and is what a
 * `class Map2DB200 : public Map2D` adapter (see include/Map2DB200.h) binds to.
 *
 *   reference (C++, Map2DFusion/)                         this library (extern "C")
 *   ----------------------------------------------------  ------------------------------------------
 *   Map2D::create(type,thread)          Map2D.cpp:51-66   m2d_create
 *   Map2D::prepare(plane,camera,frames) Map2D.h:88-89     m2d_prepare
 *     Map2DPrepare::prepare             Map2D.cpp:32-49
 *     Map2DCPUData::prepare             Map2DCPU.cpp:44-92 / MultiBandMap2DCPU.cpp:199-255
 *   Map2D::feed(img,pose)               Map2D.h:91        m2d_feed / m2d_feed_device / m2d_feed_batch
 *     Map2DCPU::renderFrame             Map2DCPU.cpp:150-336
 *     MultiBandMap2DCPU::renderFrame    MultiBandMap2DCPU.cpp:311-558
 *   Map2D::save(filename)               Map2D.h:95        m2d_save, m2d_get_image (in-memory variant)
 *     Map2DCPU::save                    Map2DCPU.cpp:523-564
 *     MultiBandMap2DCPU::save           MultiBandMap2DCPU.cpp:779-847
 *   Map2D::queueSize()                  Map2D.h:97        m2d_queue_size
 *   (raw tile state; the reference has no getter)         m2d_get_tile  (for bit-exact parity tests)
 *   legacy seam renderFramesCaller      UtilGPU.cuh:114-124  (shape of the old C++->CUDA call; NOT ported)
 *
 * Conventions: plain pointers and sizes only; no exceptions cross the ABI; every call returns an int
 * status (M2D_OK == the reference returning true, M2D_REJECTED == the reference returning false, <0 ==
 * an error the reference could not have produced, e.g. a CUDA failure). The library never calls exit().
 * Poses are 7 doubles in the reference's stream order `x y z qx qy qz qw` (GSLAM/core/SE3.h:105-117),
 * camera-to-world for frames, plane-to-world for the plane. Cameras are the 6 doubles of
 * PinHoleParameters `w h fx fy cx cy` (Map2D.h:37-43). Images are 8-bit BGR, row-major.
 */
#ifndef MAP2D_B200_H
#define MAP2D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M2D_ELE_PIXELS 256 /* Map2D.h:35  ELE_PIXELS */
#define M2D_MAX_LEVELS 9   /* BandNumber is clamped to ceil(log2(256)) = 8 -> 9 stored levels */

/* Map2D::Map2DType, Map2D.h:83. TypeGPU(2) behaves like TypeCPU(1): the reference build has no CUDA and
 * create(TypeGPU) hands back a Map2DCPU (Map2D.cpp:57-65). TypeRender(4) is the batch blender of Map2DRender.cpp: it takes m2d_render_frames, not m2d_feed*. */
enum { M2D_TYPE_NONE = 0, M2D_TYPE_CPU = 1, M2D_TYPE_GPU = 2, M2D_TYPE_MULTIBAND = 3, M2D_TYPE_RENDER = 4 };

#define M2D_OK 0
#define M2D_REJECTED 1         /* the reference's `return false` from prepare()/feed()/save() */
#define M2D_ERR_ARG (-1)
#define M2D_ERR_STATE (-2)
#define M2D_ERR_CUDA (-3)
#define M2D_ERR_NOMEM (-4)
#define M2D_ERR_UNSUPPORTED (-5)
#define M2D_ERR_IO (-6)

/* The svar keys the hot path reads, as one plain struct. */
typedef struct m2d_config {
    double scale;        /* Map2D.Scale       (Map2DCPU.cpp:75)                      default 1   */
    double resolution;   /* Map2D.Resolution  (MultiBandMap2DCPU.cpp:228), 0 = auto  default 0   */
    int weight_type;     /* Map2D.WeightType  (Map2DCPU.cpp:246)                     default 0   */
    int band_number;     /* MultiBandMap2DCPU.BandNumber (MultiBandMap2DCPU.cpp:260) default 5   */
    int force_float;     /* MultiBandMap2DCPU.ForceFloat (:444) — must be 0 (int16 path only)    */
    int background;      /* Result.BackGroundColor (:840)                            default 0   */
    int thread;          /* Map2D.Thread, informational: m2d_feed* always have the reference's thread=false
                            semantics (the call returns once the frame's bounds are decided and its work is
                            enqueued).  The thread=true behaviour -- worker thread + bounded drop-oldest queue,
                            Map2DCPU.cpp:139-142,397-413 -- is the separate m2d_ingest_* seam below, which the
                            C++ adapter (Map2DB200.h) and the Python mirror switch on when thread is set. */
    int device;          /* CUDA device ordinal                                       default 0   */
    /* Spatial tile ownership for multi-GPU runs. Tiles are keyed on ABSOLUTE tile
     * coordinates (anchored at prepare(), stable under spreadMap). owner = floor(abs / shard_span) mod
     * shard_count along shard_axis (0 = x, 1 = y). shard_count <= 1 owns everything. */
    int shard_rank, shard_count, shard_axis, shard_span;
    int collect_stats;   /* non-zero: kernels also count winners (for algorithmic-byte accounting) */
    int batch_frames;    /* frames fused per launch group in m2d_feed_batch (0 = library default) */
    int f32_mode;        /* float association of cv::pyrDown(CV_32F) on the weight pyramid: 0 = OpenCV 2.4.9 (the
                            reference's OpenCV: scalar rows, SSE columns; default), 1 = OpenCV 4.x (universal intrinsics:
                            what the cv2 4.13 in this image executes) -- only there to diff the GPU path against REAL
                            OpenCV end to end (tests/test_golden.py); differences are <= 2 ulp of a weight */
    /* TypeRender only (Map2DRender.cpp:52-310, the inline MultiBandBlender): */
    int render_blend;    /* 0 = what the reference executes: CV_32F weights, per level `if (w >= dst_w) take src` (:206-213; default);
                            1 = the disabled `#else` branch == stock OpenCV with CV_32F weights: dst += short(src * w), normalise
                            (:214-220, 264-267); 2 = the CV_16S branch == stock OpenCV with CV_16S weights: dst += short((src * w)
                            >> 8), dst_w += w, normalise (:227-249).  1 and 2 equal cv::detail::MultiBandBlender bit for bit */
    int render_bands;    /* 0 = the reference's rule: ceil(log2(sqrt(canvas area) * 5 / 100)) - 1 (:707-715); > 0 overrides */
} m2d_config;

/* Algorithmic-traffic counters, SURVEY.md §8(d). Filled by the oracle always and by the CUDA library when
 * collect_stats != 0. Index = pyramid level (weighted mode uses level 0 only). */
typedef struct m2d_stats {
    uint64_t frames_fed, frames_fused;
    uint64_t input_px;                     /* sum of W*H over fused frames                        */
    uint64_t region_px[M2D_MAX_LEVELS];    /* D_l summed over fused frames (owned tiles only)     */
    uint64_t fresh_px[M2D_MAX_LEVELS];     /* px of D_l that fell in first-touch tiles            */
    uint64_t win_px[M2D_MAX_LEVELS];       /* px of non-fresh tiles where the frame replaced state */
    uint64_t footprint_px;                 /* weighted: region px with warped alpha > 0            */
    uint64_t need_px[M2D_MAX_LEVELS];      /* multi-band, weights-first pipeline: px of the frames' Gaussian level l that had
                                              to be computed (cells a winner's Laplacian depends on); 0 from the oracle and
                                              from the dense pipeline                                */
    uint64_t needw_px[M2D_MAX_LEVELS];     /* the same for the frames' WEIGHT level l (cells in which the frame is competitive,
                                              grown by the pyrDown reach).  need_px / needw_px are also counted while
                                              m2d_profile is on (with culling active), not only with collect_stats */
} m2d_stats;

typedef struct m2d_map* m2d_handle;

void m2d_config_default(m2d_config* cfg);

/* Map2D::create — Map2D.cpp:51-66. type NONE -> M2D_ERR_UNSUPPORTED and *out = NULL. */
int m2d_create(int type, const m2d_config* cfg, m2d_handle* out);
/* Multi-GPU behind the boundary: ONE host process (the reference's host is one process: Map2D::create, Map2D.cpp:51-66), one
 * sub-map per device, tiles sharded by spatial ownership (block-cyclic strips of cfg->shard_span tiles along
 * cfg->shard_axis, on absolute tile coordinates).  The returned handle takes the same calls as a single-device one:
 * m2d_prepare / m2d_feed* (every device sees every pose and copies -- or, for device-resident frames, reads in place over
 * NVLink peer access -- only the frames under which it owns tiles; the devices fuse concurrently) / m2d_sync /
 * m2d_queue_size / m2d_reset / m2d_get_grid / m2d_last_rect / m2d_get_tile / m2d_poll_changed / m2d_get_image / m2d_save
 * (raw tiles are read to the first device over NVLink and collapsed there) / m2d_ingest_* / m2d_destroy.  Calls that only
 * make sense on one shard (set_shard, export/import, get_image_rect, state files, ...) return M2D_ERR_UNSUPPORTED.
 * n_devices == 1 is m2d_create on that device.  Requires peer access between the devices. */
int m2d_create_multi(int type, const m2d_config* cfg, int n_devices, const int* devices, m2d_handle* out);
void m2d_destroy(m2d_handle h);

/* Map2D::prepare — only the poses of the prepare-frames are used to lay out the tile grid
 * (Map2DCPU.cpp:50-61). M2D_REJECTED when the reference returns false (no frames, bad camera,
 * camera heights straddling the plane). */
int m2d_prepare(m2d_handle h, const double plane[7], const double camera[6], int n_frames,
                const double* poses /* n_frames x 7 */);

/* Map2D::feed — host image, stride in bytes. Returns after the frame's work is enqueued.  A PAGEABLE buffer
 * (cv::Mat, malloc) has been staged by the time the call returns and may be reused immediately, like the
 * reference's refcounted cv::Mat.  A PINNED buffer (m2d_alloc_host / cudaHostAlloc) is read by DMA
 * asynchronously: leave it untouched until m2d_sync() (or until m2d_queue_size() shows the frame done).
 * Batches (m2d_feed_batch, >= 8 frames) of PINNED, tightly packed frames are not staged whole: page-locked memory is
 * device-visible under UVA, so the weighted kernel samples the frames in place over PCIe, and the weights-first multi-band
 * pipeline -- which knows the winners before it needs a single image px -- pulls only the 256-byte chunks of every frame
 * that its winners' cells sample (about a third of the bytes on a 80 % / 60 % overlap survey).  Same results, bit for bit;
 * M2D_ZEROCOPY=0 in the environment restores the whole-frame staging copies.
 * A handle is not thread-safe: call it from one thread at a time (the reference serialises on its own mutex). */
int m2d_feed(m2d_handle h, const uint8_t* bgr, int w, int h_px, size_t stride, const double pose_c2w[7]);
/* Same, the image already lives in device memory (must stay valid until m2d_sync). */
int m2d_feed_device(m2d_handle h, const uint8_t* d_bgr, int w, int h_px, size_t stride,
                    const double pose_c2w[7]);
/* n frames, frame i at base + i*frame_stride; result[i] receives the per-frame status (may be NULL).
 * Semantics are exactly n sequential feed() calls; internally frames are grouped so that each map tile is
 * read and written once per group. */
int m2d_feed_batch(m2d_handle h, int n, const uint8_t* base, size_t frame_stride, int w, int h_px,
                   size_t stride, const double* poses /* n x 7 */, int on_device, int* result);

/* The same with one pointer per frame (frames that do not lie at a fixed stride: e.g. some in this GPU's memory, some mapped
 * from a neighbouring GPU's with m2d_ipc_open and sampled in place over NVLink). */
int m2d_feed_batch_ptrs(m2d_handle h, int n, const uint8_t* const* frames, int w, int h_px, size_t stride,
                        const double* poses /* n x 7 */, int on_device, int* result);

/* Sharded runs: the frames a shard does NOT need pixels for.  Applies exactly what feed() does to
 * the grid for each pose (bounds, reject test, spreadMap — Map2DCPU.cpp:163-233) and nothing else, so that every
 * shard takes the same grid decisions as an unsharded run while frame pixels travel only to the shards that own
 * tiles under them.  A pose whose footprint touches a tile this shard owns is an error (M2D_ERR_ARG, grid already
 * grown, no tile touched): the caller's delivery plan was wrong.  result[i] as for m2d_feed_batch (may be NULL). */
int m2d_feed_poses(m2d_handle h, int n, const double* poses /* n x 7 */, int* result);

/* Dry run of the bounds half of feed() (Map2DCPU.cpp:163-233) for n poses in feed order, on a COPY of the grid:
 * rects[i] = {x0,y0,x1,y1} in ABSOLUTE tile coordinates exactly as sequential feed()/m2d_feed_poses calls will
 * compute them (spreadMap growth included), or all -1 for a frame they will reject.  The map is not modified.
 * This is what a sharded host needs to decide which shard needs which frame's pixels. */
int m2d_plan_rects(m2d_handle h, int n, const double* poses /* n x 7 */, int* rects /* n x 4 */);

/* Re-partition the tile ownership of a handle that holds no tiles yet (after m2d_prepare or m2d_reset; else
 * M2D_ERR_STATE).  Same rule as m2d_config.shard_*, with strip 0 starting at absolute tile coordinate `origin`:
 * owner = floor((abs - origin) / span) mod count.  Lets the host align contiguous strips with the surveyed area,
 * which is only known once prepare() has fixed the grid. */
int m2d_set_shard(m2d_handle h, int rank, int count, int axis, int span, int origin);

int m2d_sync(m2d_handle h);
int m2d_queue_size(m2d_handle h);                 /* frames enqueued and not yet finished on the GPU */
int m2d_set_stream(m2d_handle h, void* cuda_stream /* cudaStream_t, NULL = library-owned */);
int m2d_reset(m2d_handle h);                      /* drop all tiles, keep the prepared grid */
/* Device-resident frames that are still being produced on another stream (a decoder, an NCCL receive of halo frames): the
 * NEXT m2d_feed_device / m2d_feed_batch call makes every kernel that reads frame pixels wait for `cuda_event`
 * (cudaEvent_t) first; bounds, weights and winners -- which need poses only -- start at once.  One feed call consumes it. */
int m2d_set_input_event(m2d_handle h, void* cuda_event);

/* Grid as laid out by prepare()/spreadMap (Map2DCPUData: _w,_h,_min,_max,_lengthPixel). */
int m2d_get_grid(m2d_handle h, int* w, int* h_tiles, double min_xyz[3], double max_xyz[3],
                 double* length_pixel);
/* Tile rectangle (xminInt,yminInt,xmaxInt,ymaxInt) of the most recent accepted feed (Map2DCPU.cpp:219-222),
 * in the CURRENT grid indexing. */
int m2d_last_rect(m2d_handle h, int rect[4]);
/* Raw tile state. Weighted: level must be 0, `data` receives 256*256*4 u8 BGRA, `weight` unused.
 * Multi-band: `data` receives n*n*3 int16 (interleaved BGR Laplacian, n = 256>>level), `weight` n*n f32.
 * M2D_REJECTED if the tile was never touched (or is not owned by this shard). */
int m2d_get_tile(m2d_handle h, int tx, int ty, int level, void* data, float* weight);
/* In-memory save(): bbox of touched tiles. Call with out == NULL to get the geometry, then with a buffer of
 * w*h*channels bytes. Weighted -> 4 channels BGRA (untouched tiles zero); multi-band -> 3 channels BGR,
 * collapsed, background where weight[0]==0.  While an ingest queue is open the map can grow between the two
 * calls: the second call then reads *w, *h_px, *channels (as left by the first) as the capacity of `out` and returns
 * M2D_ERR_STATE instead of overrunning it — query again. */
int m2d_get_image(m2d_handle h, uint8_t* out, int* w, int* h_px, int* channels, int* tile_min_x,
                  int* tile_min_y);
/* Map2D::save — writes the same image as a PNG (8-bit BGRA/BGR stored as RGBA/RGB). */
int m2d_save(m2d_handle h, const char* filename);

/* Final tile gather of a sharded run.  A tile travels as its raw HBM state (m2d_tile_bytes bytes,
 * library-private layout) plus its ABSOLUTE tile coordinate (stable under spreadMap and identical on every
 * shard, because every shard sees every pose).  export copies all tiles held by this handle into `dst`
 * (device or host memory) and their coordinates into abs_xy (2 ints per tile); import inserts foreign tiles
 * so that m2d_get_image()/m2d_save() on the root cover the whole map. */
size_t m2d_tile_bytes(m2d_handle h);
/* The leading bytes of a tile record that hold REFERENCE state (weighted: the BGRA tile; multi-band: per level the three
 * int16 Laplacian planes and the f32 weight plane).  What follows up to m2d_tile_bytes is library-private acceleration
 * data (per-cell lower bounds of the weight planes) that may differ between runs with identical results. */
size_t m2d_tile_state_bytes(m2d_handle h);
int m2d_tile_count(m2d_handle h);
int m2d_export_tiles(m2d_handle h, int max_tiles, int* abs_xy, uint8_t* dst, int dst_on_device, int* n_out);
int m2d_import_tiles(m2d_handle h, int n, const int* abs_xy, const uint8_t* src, int src_on_device);
/* The same for the tiles inside rect_abs = {x0,y0,x1,y1} (absolute tile coordinates, half open; NULL = all).
 * max_tiles == 0 only counts (*n_out), nothing is copied. */
int m2d_export_tiles_rect(m2d_handle h, const int* rect_abs, int max_tiles, int* abs_xy, uint8_t* dst, int dst_on_device,
                          int* n_out);
/* Give the tiles inside rect_abs back to the pool (a shard discards the halo tiles it imported for a sharded save). */
int m2d_drop_tiles_rect(m2d_handle h, const int* rect_abs, int* n_dropped);

/* Sharded save (SURVEY.md §8e "collapse sharded"; MultiBandMap2DCPU.cpp:779-847 assembles and restores ONE mosaic over the
 * bbox of all touched tiles).  m2d_tile_bbox: bbox of the tiles this handle holds, absolute tile coordinates (M2D_REJECTED if
 * none).  m2d_get_image_rect: the collapse of save() over an explicit WINDOW of tiles -- tiles the handle does not hold
 * count as the zeros the reference pastes for absent tiles -- of which only the CROP (inside the window) is written to `out`
 * (host memory, or device memory if out_on_device), crop_w x crop_h x channels bytes; out == NULL only reports the size.
 * A shard that owns tile rows [r0, r1) collapses window = global bbox columns x rows [r0 - k, r1 + k) (clipped to the
 * global bbox; the k halo rows imported raw from its neighbours, k = ceil((2^levels - 2) / 256), 1 for <= 8 levels) and
 * crops rows [r0, r1): the strips of all shards tile the reference's mosaic byte for byte, no shard ever holds the whole
 * map, and only k tile rows per boundary cross the interconnect. */
int m2d_tile_bbox(m2d_handle h, int bbox_abs[4]);
int m2d_get_image_rect(m2d_handle h, uint8_t* out, int out_on_device, const int window_abs[4], const int crop_abs[4],
                       int* w, int* h_px, int* channels);

/* Display path without GL (Map2D::draw, Map2D.h:93): the reference marks tiles `Ischanged` in renderFrame and, on the GL
 * thread, re-blends each changed tile into a texture (MultiBandMap2DCPU.cpp:702-742 -> Ele::updateTexture -> Ele::blend
 * :77-188: with HighQualityShow and all 8 neighbours present, a border of 1<<(levels-1-i) px per level is borrowed from
 * them before the restore; weighted mode uploads the BGRA tile as is, Map2DCPU.cpp:497-505).
 * m2d_poll_changed lists (and clears) the tiles touched since the last poll, in current grid coordinates;
 * m2d_get_tile_image returns one tile exactly as the reference would texture it: 256x256 BGR (multi-band) or BGRA. */
int m2d_poll_changed(m2d_handle h, int max_tiles, int* xy /* 2 ints per tile */, int* n_out);
int m2d_get_tile_image(m2d_handle h, int tx, int ty, int high_quality, uint8_t* out, int* channels);

/* Map2DRender (Map2D::TypeRender = 4) -- Map2DFusion/Map2DRender.cpp.  The reference's feed() only queues (thread = true) or
 * returns false (thread = false, renderFrame :464-467); its worker takes EVERYTHING queued -- the prepare-frames first -- as one
 * batch through renderFrames (:479-760): every frame is warped into its own bounding box (8UC3 bilinear BORDER_REFLECT + the
 * 8-bit weight image, nearest), the map is spread, and an inline copy of cv::detail::MultiBandBlender (:52-310) blends the
 * batch onto a canvas of whole tiles with ceil(log2(sqrt(area) * 0.05)) - 1 bands (m2d_config.render_bands overrides); the
 * result goes to "result.jpg" and the thread stops.  m2d_render_frames is that batch call: frame i at base + i * frame_stride
 * (host or device memory), poses camera-to-world.  result[i] = M2D_OK, or M2D_REJECTED for a frame the reference skips
 * (oblique view).  Not built: the GUI (cv::imshow / waitKey of every pyramid level, Map2DRender.ShowPyrLaplace) and the seam
 * finder (Map2DRender.EnableSeam, cv::detail::DpSeamFinder: a sequential dynamic program) -- i.e. EnableSeam = 0.
 * The blended canvas stays in the handle: m2d_get_image returns it as 8-bit BGR (what cv::imwrite stores), m2d_save writes a
 * PNG instead of a JPEG, m2d_render_get returns the blender's raw outputs: the CV_16SC3 result (w*h*3 int16, masked px 0), its
 * mask (w*h, 255 where the level-0 weight exceeds WEIGHT_EPS), the band count and the ABSOLUTE tile coordinate of the canvas
 * origin.  m2d_feed* on a TypeRender handle return M2D_REJECTED like the reference. */
int m2d_render_frames(m2d_handle h, int n, const uint8_t* base, size_t frame_stride, int w, int h_px, size_t stride,
                      const double* poses /* n x 7 */, int on_device, int* result);
int m2d_render_get(m2d_handle h, int16_t* result16, uint8_t* mask, int* w, int* h_px, int* num_bands, int* tile_x0, int* tile_y0);

/* Checkpoint / resume of the mosaic (the reference can only save the final PNG; its SLAM map has
 * MapHash::save/load, GSLAM-DIYSLAM/src/zhaoyong/MapHash.cpp:376,458).  save_state writes the prepared grid
 * (camera, plane, extents, origin) and the raw state of every tile held by this handle; load_state restores them
 * into a handle created with the same type and band number, after which feeding continues exactly as if it had
 * never been interrupted (bit-identical results).  File format: private, versioned, little-endian. */
int m2d_save_state(m2d_handle h, const char* filename);
int m2d_load_state(m2d_handle h, const char* filename);

int m2d_get_stats(m2d_handle h, m2d_stats* out);
const char* m2d_last_error(m2d_handle h);
/* Number of CUDA kernels this handle has launched so far (bench.py's gpu_launches). */
uint64_t m2d_launch_count(m2d_handle h);

/* Per-kernel-class device timing: while enabled, every launch on the handle's stream is bracketed by CUDA
 * events; m2d_get_kernel_times synchronises and returns accumulated milliseconds and launch counts per class
 * (and clears them).  bench.py uses it for the live roofline measurement; leave it off otherwise. */
#define M2D_KERNEL_CLASSES 16
enum { M2D_K_WEIGHTED = 0, M2D_K_MB_WARP = 1, M2D_K_MB_PYRDOWN = 2, M2D_K_MB_SELECT = 3, M2D_K_COLLAPSE = 4, M2D_K_MISC = 5, M2D_K_MB_PYRTAIL = 6,
       /* weights-first multi-band pipeline (default): weight warp / weight pyramid / decide / propagate (weight + image side) /
        * sparse image warp / sparse image pyramid / winners' Laplacian / competitive-cell bounds */
       M2D_K_MBW_WARP = 7, M2D_K_MBW_PYR = 8, M2D_K_MBS_DECIDE = 9, M2D_K_MBS_PROPAGATE = 10, M2D_K_MBS_WARP = 11, M2D_K_MBS_PYR = 12,
       M2D_K_MBS_LAP = 13, M2D_K_MBC_BOUNDS = 14,
       M2D_K_RENDER = 15 /* TypeRender: warp / pyramid / blend / normalise / final of m2d_render_frames */ };
int m2d_profile(m2d_handle h, int enable);
int m2d_get_kernel_times(m2d_handle h, double* ms /* M2D_KERNEL_CLASSES */, uint64_t* count /* M2D_KERNEL_CLASSES */);

/* Ingest seam — the step in front of feed().  In the reference the tracker thread converts each
 * frame BGRA -> BGR (GSLAM-DIYSLAM/src/zhaoyong/TrackerOpt.cpp:374-383) and hands (image, pose) over through a bounded
 * queue that DROPS THE OLDEST entry when full (src/DataTrans.h:54-68, capacity 30; Map2DCPU's own thread=true queue does
 * the same with capacity 20, Map2DCPU.cpp:139-142,397-413).  m2d_ingest_open (after m2d_prepare) allocates a ring of
 * pinned BGR8 slots and starts one worker thread that pops the queue in order and feeds it in grouped launches.
 * m2d_ingest_push copies (channels 3 = BGR) or converts (channels 4 = BGRA) the caller's pixels into a slot and
 * returns at once — the caller's buffer is free again; it never blocks, a full queue loses its oldest frame.  Returns
 * M2D_REJECTED for a frame whose size is not the camera's.  push may be called from any thread; while ingest is open
 * every other call on the handle is serialised with the worker.  m2d_queue_size counts queued frames too.
 * pause(1) holds the worker (frames keep queueing/dropping), drain blocks until everything queued has been fused,
 * close drains, joins the worker and frees the ring (m2d_destroy closes implicitly). */
int m2d_ingest_open(m2d_handle h, int capacity, int start_paused);
/* The same with room for `seed_frames` frames queued up front WITHOUT dropping: in the reference Map2DPrepare::_frames is
 * both the prepare set and the worker's queue (Map2D.cpp:42), so with thread=true the worker first renders the
 * prepare-frames, however many there are; only a later feed() drops (one) oldest entry when the queue holds more than
 * `capacity` (Map2DCPU.cpp:141-142).  Open paused, push the prepare-frames, then m2d_ingest_pause(h, 0). */
int m2d_ingest_open_seeded(m2d_handle h, int capacity, int seed_frames, int start_paused);
int m2d_ingest_push(m2d_handle h, const uint8_t* pixels, int w, int h_px, size_t stride, int channels,
                    const double pose_c2w[7]);
int m2d_ingest_pause(m2d_handle h, int paused);
int m2d_ingest_drain(m2d_handle h);
int m2d_ingest_close(m2d_handle h);
/* close WITHOUT draining: frames still queued are discarded (counted as dropped).  What a second prepare() does to the
 * old queue in the reference: the new Map2DPrepare object brings its own frame deque (Map2DCPU.cpp:105-125). */
int m2d_ingest_abort(m2d_handle h);
int m2d_ingest_stats(m2d_handle h, uint64_t* pushed, uint64_t* dropped, uint64_t* fed, uint64_t* fused);

/* Map2DUpdate — the Google-map overlay command of the display loop (MultiBandMap2DCPU.cpp:744-757, consumed by
 * Map2DItem.cpp:36-99): GPS corners of tile (tx, ty).  Pure host arithmetic, no handle: the tile's ground corners are
 * formed in FLOAT like the reference (`float x0 = min.x + x*eleSize`, `float x1 = x0 + eleSize`), moved to the world
 * frame with plane * (x, y, 0) and converted by pi::calcLngLatFromDistance (PIL/src/hardware/Gps/utils_GPS.cpp:133-160,
 * WGS-84 units of longitude/latitude at the origin's latitude).  gps_origin = {lng, lat} (svar "GPS.Origin");
 * tl/br receive {lng, lat, 0}.  The texture that goes with it is m2d_get_tile_image (+ level-0 weights as alpha). */
int m2d_tile_gps_corners(const double plane[7], double grid_min_x, double grid_min_y, double ele_size, int tx, int ty,
                         const double gps_origin[2], double tl[3], double br[3]);

/* Introspection for tests: the dependency-reach table of the weights-first multi-band pipeline (kernels.cu
 * make_reach_table).  lo/hi[m*6 + k] = how many 32-px cells below/above the cell of a level-m winner the Gaussian level
 * k must be valid (255 = no dependency).  levels = band_number + 1 <= 6.  Pure host arithmetic, no handle. */
int m2d_reach_table(int levels, unsigned char lo[36], unsigned char hi[36]);
/* The same for the WEIGHT side (kernels_wf.cu make_weight_reach_table): cells of weight level k that a frame's competitive
 * cell of level m depends on (the pyrDown chain only). */
int m2d_weight_reach_table(int levels, unsigned char lo[36], unsigned char hi[36]);
/* Introspection for tests: the closed-form bounds the weights-first pipeline culls with (csrc/bounds.h, the very code the
 * bounds kernel runs).  hinv = inverse homography of a frame (region px -> source px, as m2d_compute_bounds returns it;
 * rounded to float like the kernel's copy), (nx, ny) its region in tiles, (sw, sh) the source size; on return
 * lo <= W_level(u) <= hi for every px u of cell (cx, cy) (32 x 32 level-0 px of the region) of the frame's weight
 * pyramid (MultiBandMap2DCPU.cpp:449-474).  Pure host arithmetic, no handle. */
int m2d_cell_weight_bounds(const double hinv[9], int nx, int ny, int sw, int sh, int weight_type, int level, int cx,
                           int cy, float* lo, float* hi);

/* Introspection for tests: the source rectangle pull mode fetches for one 32 x 32 level-0 cell of a frame's region (csrc/bounds.h
 * pull_cell_rect, the very code mbs_mark runs).  (X0, Y0) = region px of the cell's corner, hinv as m2d_compute_bounds returns
 * it; rect = {lox, hix, loy, hiy}, inclusive source px, containing every tap the image warp of the cell can read (bilinear,
 * 1/32-px rounding, BORDER_REFLECT).  M2D_REJECTED: degenerate geometry, rect is the whole frame.  Pure host arithmetic. */
int m2d_pull_cell_rect(const double hinv[9], int X0, int Y0, int sw, int sh, int rect[4]);

/* Frame buffers that the other PROCESSES of a multi-GPU job can map (CUDA IPC).  One process per GPU keeps the frames its own
 * flight lines produced in a buffer from m2d_device_alloc, exports it once (m2d_ipc_export -> 64 opaque bytes, sent to the
 * neighbours by any means), and a neighbour that owns tiles under some of those frames maps the buffer (m2d_ipc_open, peer
 * access over NVLink/NVSwitch is enabled on the way) and passes the mapped addresses to m2d_feed_batch_ptrs: its kernels
 * sample the remote frames in place, so only the px that are really needed cross the link and no halo copy is ever made. */
void* m2d_device_alloc(int device, size_t bytes);
void m2d_device_free(int device, void* p);
int m2d_ipc_export(void* dptr, unsigned char handle[64]);
int m2d_ipc_open(int device, const unsigned char handle[64], void** dptr);
int m2d_ipc_close(int device, void* dptr);

/* Pinned host staging helpers for callers that want truly asynchronous m2d_feed(). */
void* m2d_alloc_host(size_t bytes);
void m2d_free_host(void* p);

/* The tile-overlap/bounds kernel on its own (SURVEY.md §8a A4+A7): for n poses against the CURRENT grid,
 * rect[i] = {xminInt,yminInt,xmaxInt,ymaxInt} (or all -1 when the frame is rejected) and hinv[i] = the
 * inverse homography (region px -> source px, row-major 3x3). No spreadMap is applied: out-of-grid frames
 * report indices outside [0,w]x[0,h].  NOT on the feed path: m2d_feed* take these decisions one frame after the
 * other on the host (the same geom.h code; a frame's grid depends on the spreadMap of the frames before it), so this
 * kernel is a batch query for planners and for the host/device bit-equality test of geom.h. */
int m2d_compute_bounds(m2d_handle h, int n, const double* poses, int* rects /* n x 4 */,
                       double* hinv /* n x 9 */);

#ifdef __cplusplus
}
#endif
#endif /* MAP2D_B200_H */
