// kernels.cuh — launch parameter blocks and host-callable launchers of the sm_100a kernels.
//
// Execution model: frames are fused in GROUPS of up to K frames.  All per-frame stages of a group
// (warp, pyramid) run as single launches over (work, frame); the order-dependent stage (select) is
// TILE-CENTRIC: one work item per touched tile, looping over the group's frames that cover it (feed order for
// multi-band, best-first for weighted -- both rules are order-free once the frame index is carried along), so every
// map tile is read and written once per group and the result equals K sequential feed() calls.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/map2d_b200.h"
#include "geom.h"

namespace m2d {

constexpr int kEle = M2D_ELE_PIXELS;

// One frame of a group, as the kernels see it.  Lives in a device array (uploaded once per group).
struct FrameJob {
    double hinv[9];            // region px -> source px (inverse homography, cv::warpPerspective convention)
    float hinvf[9];            // the same in FP32: conservative culling only, never for sampling
    const uint8_t* raw;        // BGR8 source in HBM, sampled in place (3-byte taps via aligned word loads)
    int raw_stride;
    int nx, ny;                // frame region in tiles
    int wx, wy, wnx, wny;      // pyramid window inside the region (tiles, relative to the region origin)
    unsigned long long g_off[M2D_MAX_LEVELS];  // byte offset in the group scratch of level l's u8x4 Gaussian
    unsigned long long w_off[M2D_MAX_LEVELS];  // ... of level l's f32 weight plane
    const uint8_t* pull_src;   // pull mode (pinned host frames, weights-first pipeline): device-visible address of the caller's
                               // frame in HOST memory; `raw` is then a staging slot in HBM that mbs_pull fills with the needed
                               // 256-byte chunks only (NULL otherwise)
};

// Tile-centric work list of a group.
struct TileWork {
    uint8_t* state;            // tile state in HBM
    int first, count;          // entries [first, first+count): feed order (multi-band) / best-first (weighted)
    int fresh;                 // tile allocated by this group: no state to read
};
struct TileEntry {
    int frame;                 // index into the group's FrameJob array == feed order inside the group
    short rtx, rty;            // tile position inside that frame's region
};

struct EntryRef {                // per (tile entry, level): where the frame's weight plane starts under the tile (kernels_wf.cu)
    unsigned long long base;     // address of the weight of the tile's px (0, 0) at that level
    int stride;                  // row pitch of the plane in px (window width in tiles * tile side)
    int cell0;                   // index into the per-frame cell flag arrays of the tile's cell (0, 0) at that level
};

struct GroupParams {
    const FrameJob* jobs;
    const TileWork* tiles;
    const TileEntry* entries;
    int n_frames, n_tiles;
    int sw, sh;                // source frame size
    int levels;                // multi-band: band_num + 1
    int weight_type;           // Map2D.WeightType (alpha = dis or dis^2)
    int f32_mode;              // m2d_config.f32_mode: float association of the weight pyrDown (0 = OpenCV 2.4.9, 1 = 4.x)
    const uint8_t* alpha;      // weighted: sw*sh alpha image (Map2DCPU.cpp:236-258)
    const float* wimg;         // multi-band: sw*sh float weight image (MultiBandMap2DCPU.cpp:396-418)
    uint8_t* scratch;          // multi-band: group scratch pyramid
    unsigned long long* stats; // optional counters (collect_stats)
    unsigned long long* need_stats;  // optional: [20 + k] px of Gaussian level k computed, [26 + k] px of weight level k computed
    int max_wnx, max_wny;      // largest pyramid window of the group (tiles): grid bounds of per-frame kernels
    // weights-first multi-band (kernels_wf.cu): per (frame, level) cell flags, work lists and the winner map.
    // A CELL is 32 x 32 level-0 px of a frame's window; at level l it is (32 >> l)^2 px (levels <= 6).
    uint8_t* comp;             // [n_frames][levels][cells_max]: the frame is COMPETITIVE in that cell at that level (its
                               // weight upper bound reaches the best lower bound of the other frames / the tile state)
    uint8_t* needw;            // same shape: the frame's WEIGHT level must be valid in that cell
    uint8_t* win;              // same shape: the frame wins a px of that cell at that level
    uint8_t* need;             // same shape: the frame's Gaussian level must be valid in that cell
    int cells_max;             // (max_wnx * 8) * (max_wny * 8)
    uint32_t* cmask;           // [n_tiles][levels][64 cells][mask_words]: bit i = entry i of the tile's list is competitive
    int mask_words;            // ceil(max entries per tile / 32)
    uint32_t* lists;           // [2 (0 = weight, 1 = image)][levels][list_cap] work items (frame << 16 | cell) ...
    unsigned* list_count;      // ... and their lengths [2][levels] (device counters)
    int list_cap;              // n_frames * cells_max
    uint16_t* wmap;            // [n_tiles][wmap_stride]: winning entry per pyramid px of the tile (0xFFFF = none)
    int wmap_stride;           // >= TileLayout::px_off[levels], even
    unsigned char reach_lo[6][6], reach_hi[6][6];    // image: [win level m][Gaussian level k] in cells, 0xFF = none (make_reach_table)
    unsigned char wreach_lo[6][6], wreach_hi[6][6];  // weights: [competitive level m][weight level k]   (make_weight_reach_table)
    EntryRef* etable;          // [n_entries][levels]
    int use_tma;               // pyrDown 0 -> 1 through TMA-staged shared-memory patches (M2D_TMA=0: register-window kernel)
    int cull;                  // 0: every entry of a tile is treated as competitive (collect_stats, M2D_WCULL=0)
    // pull mode: which 256-byte chunks of every frame the image stage will read (bit per chunk, src_words words per frame)
    uint32_t* src_bits;
    int src_words;
    unsigned long long frame_bytes;
};

// Tile state layout in HBM (multi-band): for each level l (side n = 256>>l): B,G,R int16 planes then f32 weight.
struct TileLayout {
    int levels;
    size_t lap_off[M2D_MAX_LEVELS];
    size_t wgt_off[M2D_MAX_LEVELS];
    size_t cmin_off;           // f32 [levels][64]: lower bound of the tile's weight over each 32-px cell of each level (never
                               // above the true minimum; weights only grow, so a stale value stays valid)
    size_t bytes;
    int px_off[M2D_MAX_LEVELS + 1];
};
TileLayout make_tile_layout(int levels);

struct MosaicLevel {
    int16_t* g[3];
    int w, h;
};
struct MosaicSet {
    MosaicLevel lv[M2D_MAX_LEVELS];
    float* w0;                 // level-0 weight mosaic (background mask)
};
struct SubPaste {               // a rectangle of one level of one tile -> a position in the bordered tile pyramid
    const uint8_t* tile;
    int level, sx, sy, w, h, dx, dy;
};
struct PasteItem {
    const uint8_t* tile;
    int tx, ty;                // tile position inside the mosaic
};

// ---- Map2DRender (type 4): the batch blender, Map2DRender.cpp:52-310 + 479-760 (kernels_render.cu) ----
constexpr int kRenderMaxLevels = 16;
struct RenderJob {               // one frame of a render batch
    double hinv[9];              // px of the frame's own warped image -> source px (Map2DRender.cpp:579-586)
    const uint8_t* raw;          // BGR8 source, sampled in place
    int raw_stride;
    int iw, ih;                  // size of the warped image (sizes[idx], :572)
    int left, top;               // where it sits inside the bordered sub-image (copyMakeBorder, :147-151)
    int x_tl, y_tl;              // the sub-image on the canvas, level 0 (multiples of 1 << num_bands, :118-139)
    int sw, sh;                  // sub-image size, level 0 (multiples of 1 << num_bands)
    unsigned long long g_off[kRenderMaxLevels];  // byte offsets in the batch scratch: packed u8x4 Gaussian of level l ...
    unsigned long long w_off[kRenderMaxLevels];  // ... and its weight plane (f32, or s16 for render_blend 2)
};
struct RenderCanvas {            // dst_pyr_laplace_ / dst_band_weights_ (:87-100)
    int levels;                  // num_bands + 1
    int w[kRenderMaxLevels], h[kRenderMaxLevels];
    int16_t* lap[kRenderMaxLevels][3];   // planar B, G, R
    void* wgt[kRenderMaxLevels];         // f32, or s16 for render_blend 2
};
cudaError_t launch_rnd_weight_image(int sw, int sh, uint8_t* out, cudaStream_t stream);
cudaError_t launch_rnd_warp(const RenderJob* d_jobs, int n_frames, int max_px, uint8_t* scratch, const uint8_t* wimg, int src_w, int src_h,
                            int s16_weights, cudaStream_t stream);
cudaError_t launch_rnd_pyrdown(const RenderJob* d_jobs, int n_frames, int max_px_l0, uint8_t* scratch, int level, int f32_mode, int s16_weights,
                               cudaStream_t stream);
cudaError_t launch_rnd_blend(const RenderJob* d_jobs, int n_frames, const uint8_t* scratch, const RenderCanvas& C, int blend, cudaStream_t stream);
cudaError_t launch_rnd_normalize(const RenderCanvas& C, int blend, cudaStream_t stream);
cudaError_t launch_rnd_final(const RenderCanvas& C, int s16_weights, int wf, int hf, int16_t* out16, uint8_t* out8, uint8_t* mask, cudaStream_t stream);

// launchers (asynchronous on `stream`; return the launch error)
cudaError_t launch_weight_images(int sw, int sh, int weight_type, uint8_t* alpha, float* wimg, cudaStream_t stream);
cudaError_t launch_bounds(const GridGeom& g, int n, const double* d_poses, FrameBounds* d_out, cudaStream_t stream);
cudaError_t launch_weighted_group(const GroupParams& p, cudaStream_t stream);
cudaError_t launch_mb_warp(const GroupParams& p, cudaStream_t stream);
cudaError_t launch_mb_pyrdown(const GroupParams& p, int level /* src level */, cudaStream_t stream);
cudaError_t launch_mb_pyrtail(const GroupParams& p, int l_first, cudaStream_t stream);
cudaError_t launch_mb_select(const GroupParams& p, const TileLayout& lay, cudaStream_t stream);
// weights-first multi-band pipeline (kernels_wf.cu)
cudaError_t launch_mbc_bounds(const GroupParams& p, const TileLayout& lay, cudaStream_t stream);          // competitive cells
cudaError_t launch_mbc_entry_table(const GroupParams& p, int n_entries, cudaStream_t stream);
cudaError_t launch_mbx_propagate(const GroupParams& p, int image, cudaStream_t stream);                   // need flags + work lists
cudaError_t launch_mbw_warp(const GroupParams& p, int ctas, cudaStream_t stream);                         // level-0 weights, listed cells
cudaError_t launch_mbx_pyrdown(const GroupParams& p, int image, int level, int ctas, cudaStream_t stream); // l -> l+1, listed cells
cudaError_t launch_mbs_decide(const GroupParams& p, const TileLayout& lay, cudaStream_t stream);
cudaError_t launch_mbs_warp(const GroupParams& p, int ctas, cudaStream_t stream);                         // level-0 Gaussian, listed cells
cudaError_t launch_mbs_lap(const GroupParams& p, const TileLayout& lay, cudaStream_t stream);
cudaError_t launch_mbs_mark(const GroupParams& p, int ctas, cudaStream_t stream);   // pull mode: chunks of the source frames the listed cells sample
cudaError_t launch_mbs_pull(const GroupParams& p, int ctas, cudaStream_t stream);   // ... copied from pinned host memory into the staging slots
void make_reach_table(int levels, unsigned char lo_tab[6][6], unsigned char hi_tab[6][6]);
void make_weight_reach_table(int levels, unsigned char lo_tab[6][6], unsigned char hi_tab[6][6]);
cudaError_t launch_tile_copy(uint8_t* const* d_tiles, int n, uint8_t* buf, size_t tile_bytes, int to_buf, cudaStream_t stream);
cudaError_t launch_bgra_paste(const PasteItem* d_items, int n_items, uint32_t* mosaic, int mosaic_w, cudaStream_t stream);
cudaError_t launch_mosaic_paste(const PasteItem* d_items, int n_items, const TileLayout& lay, const MosaicSet& ms, cudaStream_t stream);
cudaError_t launch_sub_paste(const SubPaste* d_items, int n_items, const TileLayout& lay, const MosaicSet& ms, cudaStream_t stream);
cudaError_t launch_tile_crop(MosaicLevel m0, int border, const float* w0, uint8_t* out, cudaStream_t stream);
cudaError_t launch_mosaic_upadd(MosaicLevel coarse, MosaicLevel fine, cudaStream_t stream);
cudaError_t launch_mosaic_final(MosaicLevel m0, const float* w0, int background, uint8_t* out_bgr, cudaStream_t stream);

}  // namespace m2d
