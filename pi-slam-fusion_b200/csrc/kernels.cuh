// kernels.cuh — launch parameter blocks and host-callable launchers of the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/map2d_b200.h"
#include "geom.h"

namespace m2d {

constexpr int kEle = M2D_ELE_PIXELS;
constexpr int kMaxRectTiles = 1024;  // fresh-tile bitmask capacity per frame (32 x 32 tiles = 67 Mpx region)

// Geometry of one frame against the tile grid, as the kernels need it.
struct FrameRect {
    int rx0, ry0, nx, ny;     // frame region: origin tile (grid indexing) and size in tiles
    int wx0, wy0, wnx, wny;   // window = bbox of the tiles this shard owns inside the region (grid indexing)
};

// ------------------------------------------------------------------ weighted mode (Map2DCPU)
struct WeightedParams {
    double hinv[9];               // region px -> source px
    const uint8_t* src;           // BGR8 frame in HBM
    int src_stride, sw, sh;
    const uint8_t* alpha;         // sw*sh distance-to-centre alpha (Map2DCPU.cpp:236-258), built once per size
    uint8_t* const* table;        // device tile table: pointer per grid slot (NULL = none / not owned)
    int grid_w;
    FrameRect r;
    uint32_t fresh[kMaxRectTiles / 32];  // bit (ty-ry0)*nx+(tx-rx0): tile allocated by this frame (skip the read)
    unsigned long long* stats;    // optional: [0] footprint px, [1] wins on non-fresh tiles
};

// ------------------------------------------------------------------ multi-band mode (MultiBandMap2DCPU)
// Per-frame scratch pyramid over the window (planar: 3 x int16 Gaussian planes + 1 x f32 weight plane).
struct PyrLevel {
    int16_t* g[3];
    float* w;
    int ww, wh;   // window size at this level (px)
    int rw, rh;   // full frame-region size at this level (px): borders reflect HERE, not at the window edge
    int ox, oy;   // window origin inside the region at this level (px)
};

struct MultibandParams {
    double hinv[9];
    const uint8_t* src;
    int src_stride, sw, sh;
    const float* wimg;            // sw*sh float weight image (MultiBandMap2DCPU.cpp:396-418)
    uint8_t* const* table;
    int grid_w;
    FrameRect r;
    uint32_t fresh[kMaxRectTiles / 32];
    int levels;                   // band_num + 1
    PyrLevel lv[M2D_MAX_LEVELS];
    unsigned long long* stats;    // optional: [l] wins on non-fresh tiles at level l
};

// Tile state layout in HBM (multi-band): for each level l (side n = 256>>l): B,G,R int16 planes then f32 weight.
struct TileLayout {
    int levels;
    size_t lap_off[M2D_MAX_LEVELS];  // byte offset of the first int16 plane of level l
    size_t wgt_off[M2D_MAX_LEVELS];  // byte offset of the f32 weight plane of level l
    size_t bytes;
    int px_off[M2D_MAX_LEVELS + 1];  // cumulative pixel count (for flat work decomposition)
};
TileLayout make_tile_layout(int levels);

// Collapse (save / get_image): per-level mosaics over the bbox of touched tiles.
struct MosaicLevel {
    int16_t* g[3];
    int w, h;
};

// launchers (all asynchronous on `stream`; return the cudaError of the launch)
cudaError_t launch_weight_images(int sw, int sh, int weight_type, uint8_t* alpha, float* wimg, cudaStream_t stream);
cudaError_t launch_bounds(const GridGeom& g, int n, const double* d_poses, FrameBounds* d_out, cudaStream_t stream);
cudaError_t launch_weighted(const WeightedParams& p, cudaStream_t stream);
cudaError_t launch_mb_warp(const MultibandParams& p, cudaStream_t stream);
cudaError_t launch_mb_pyrdown(const MultibandParams& p, int level /* src level */, cudaStream_t stream);
cudaError_t launch_mb_select(const MultibandParams& p, const TileLayout& lay, cudaStream_t stream);
cudaError_t launch_mosaic_clear(MosaicLevel m, float* w0, cudaStream_t stream);
cudaError_t launch_mosaic_paste(const uint8_t* tile, const TileLayout& lay, int level, MosaicLevel m, float* w0,
                                int tx, int ty, cudaStream_t stream);
cudaError_t launch_mosaic_upadd(MosaicLevel coarse, MosaicLevel fine, cudaStream_t stream);
cudaError_t launch_mosaic_final(MosaicLevel m0, const float* w0, int background, uint8_t* out_bgr, cudaStream_t stream);

}  // namespace m2d
