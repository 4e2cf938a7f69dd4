// kernels_wf.cu — the WEIGHTS-FIRST multi-band pipeline, default for groups of >= 4 frames, <= 6 levels.
//
// The weight pyramids depend on geometry only, so the winner of every pyramid px (MultiBandMap2DCPU.cpp:539-547: the
// LAST frame whose weight is >= everything before it) is known before a single image sample is taken, and most of the
// (frame, px) pairs can be ruled out before a single WEIGHT is computed.  Stages of one group:
//
//   0. mbc_bounds     (tile-centric)  closed-form bounds of every covering frame's weight over every 32-px cell of
//                                     every level (bounds.h); a frame is COMPETITIVE in a cell when its upper bound
//                                     reaches the best lower bound of the other frames / of the tile state
//   1. mbx_propagate  (weights)       cells in which a frame's weight level k must be valid (competitive cells of the
//                                     levels >= k, grown by the reach of the pyrDown taps) -> flags + work lists
//      mbw_warp, mbx_pyrdown<W>       weights (f32, nearest, 2.4.9 float filter) in the listed cells only
//   2. mbs_decide     (tile-centric)  arg-max over the competitive frames -> tile weights, winner map, `win` flags
//   3. mbx_propagate  (image)         cells in which a frame's Gaussian level k is needed by one of its winners
//   4. mbs_warp, mbx_pyrdown<G>       image (u8x4, 1/32-px bilinear, BORDER_REFLECT; packed integer filter), listed cells
//   5. mbs_lap        (tile-centric)  winners only: G_l - pyrUp(G_{l+1}) -> Laplacian planes of the tile
//
// A CELL is 32 x 32 level-0 px of a frame's window; at level l it is (32 >> l)^2 px (levels <= 6), so one cell grid
// (8 x 8 cells per tile) serves every level: cell of level-l px p = (p << l) >> 5.  The per-frame stages are PERSISTENT
// kernels over compacted work lists (one CTA slot per SM x residency, grid-stride over (frame, cell) items), so an
// empty cell costs nothing.  Px outside the listed cells are never read by a listed px: they may hold anything.
// Results are bit-identical to the dense pipeline (kernels.cu) and to the oracle.
#include "bounds.h"
#include "device_common.cuh"

namespace m2d {

// ---------------------------------------------------------------------------------------------------------
// cell-aligned mapping of the tile-centric kernels: blockIdx.y, threadIdx.x -> (level, cell of the tile, px)
//   level 0: 64 CTAs, one 32x32-px cell each (16 x 16 quads)      level 3: 1 CTA, 64 cells of 2 x 2 quads
//   level 1: 16 CTAs, 4 cells of 8 x 8 quads                      level 4 + 5: 1 CTA: 64 quads (one per cell) of level 4,
//   level 2: 4 CTAs, 16 cells of 4 x 4 quads                                   64 single px (one per cell) of level 5
// A thread owns a 2x2 quad (levels 0..4) or one px (level 5); all its px lie in ONE cell.
// ---------------------------------------------------------------------------------------------------------
struct CellMap { int l, n, cell, px, py, npx; bool valid; };

__host__ __device__ inline int cellmap_ctas(int levels) {
    const int per_level[6] = {64, 16, 4, 1, 1, 0};   // level 5 shares the CTA of level 4
    int c = 0;
    for (int l = 0; l < levels && l < 6; l++) c += per_level[l];
    return c;
}
__device__ __forceinline__ CellMap cell_map(int by, int t, int levels) {
    CellMap m;
    m.valid = true; m.npx = 4;
    int tt;
    if (by < 64) { m.l = 0; m.cell = by; tt = t; }
    else if (by < 80) { m.l = 1; m.cell = (by - 64) * 4 + (t >> 6); tt = t & 63; }
    else if (by < 84) { m.l = 2; m.cell = (by - 80) * 16 + (t >> 4); tt = t & 15; }
    else if (by == 84) { m.l = 3; m.cell = t >> 2; tt = t & 3; }
    else {
        if (t < 64) { m.l = 4; m.cell = t; tt = 0; }
        else { m.l = 5; m.cell = t - 64; tt = 0; m.npx = 1; m.valid = t < 128 && levels > 5; }
    }
    m.n = kEle >> m.l;
    const int side = m.n >> 3;                 // cell side in px: 32, 16, 8, 4, 2, 1
    const int Q = max(side >> 1, 1);           // quads per cell side
    const int qy = tt / Q, qx = tt - qy * Q;
    m.px = (m.cell & 7) * side + 2 * qx;
    m.py = (m.cell >> 3) * side + 2 * qy;
    if (m.l >= levels) m.valid = false;
    return m;
}

// ---------------------------------------------------------------------------------------------------------
// 0. competitive cells.  One CTA per touched tile, one thread per (level, cell of the tile).  Pass 1: the best lower
// bound L over the tile's entries and the tile state (cmin).  Pass 2: entry i is competitive iff hi_i >= L ('>=': ties
// are decided by feed order, so an equal frame may still win; where everything is 0 -- the mosaic's outer rim -- every
// covering frame stays in, because `0 >= 0` lets the last one overwrite).  A frame that is NOT competitive is strictly
// below the frame (or state) that attains L at every px of the cell: it can never be the final winner there.
// Outputs: the per-(tile, level, cell) bit mask over the tile's entries, the per-frame `comp` cell flags, and cmin reset
// to +inf in every cell the decide stage will revisit (it re-establishes the exact minimum there).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(384) mbc_bounds_kernel(const __grid_constant__ GroupParams p, const __grid_constant__ TileLayout lay) {
    const TileWork T = p.tiles[blockIdx.x];
    const int l = threadIdx.x >> 6, cell = threadIdx.x & 63;
    if (l >= p.levels) return;
    const int ccx = cell & 7, ccy = cell >> 3;
    float* cmin = reinterpret_cast<float*>(T.state + lay.cmin_off) + l * 64 + cell;
    uint32_t* cm = p.cmask + (((size_t)blockIdx.x * p.levels + l) * 64 + cell) * p.mask_words;
    float best_lo = -INFINITY;
    constexpr int kCache = 48;     // upper bounds of the first entries are kept (thread-local memory) instead of recomputed
    float hic[kCache];
    if (p.cull) {
        if (!T.fresh) best_lo = *cmin;
        for (int i = 0; i < T.count; i++) {
            const TileEntry E = p.entries[T.first + i];
            const FrameJob& J = p.jobs[E.frame];
            float lo, hi;
            cell_weight_bounds(J.hinvf, J.nx, J.ny, p.sw, p.sh, p.weight_type, l, E.rtx * 8 + ccx, E.rty * 8 + ccy, &lo, &hi);
            best_lo = fmaxf(best_lo, lo);
            if (i < kCache) hic[i] = hi;
        }
    }
    uint32_t any = 0u;
    for (int w = 0; w < p.mask_words; w++) {
        uint32_t bits = 0u;
        const int i1 = min(T.count, (w + 1) * 32);
        for (int i = w * 32; i < i1; i++) {
            const TileEntry E = p.entries[T.first + i];
            const FrameJob& J = p.jobs[E.frame];
            bool in = true;
            if (p.cull) {
                float lo, hi;
                if (i < kCache) hi = hic[i];
                else cell_weight_bounds(J.hinvf, J.nx, J.ny, p.sw, p.sh, p.weight_type, l, E.rtx * 8 + ccx, E.rty * 8 + ccy, &lo, &hi);
                in = !(hi < best_lo);   // NaN-safe: a bound that could not be evaluated never culls
            }
            if (in) {
                bits |= 1u << (i & 31);
                p.comp[cell_base(p, E.frame, l) + (size_t)((E.rty - J.wy) * 8 + ccy) * (J.wnx * 8) + ((E.rtx - J.wx) * 8 + ccx)] = 1;
            }
        }
        cm[w] = bits;
        any |= bits;
    }
    if (any) *cmin = INFINITY;
}
cudaError_t launch_mbc_bounds(const GroupParams& p, const TileLayout& lay, cudaStream_t stream) {
    if (p.n_tiles == 0) return cudaSuccess;
    mbc_bounds_kernel<<<p.n_tiles, 64 * min(p.levels, 6), 0, stream>>>(p, lay);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// 1./3. propagate: level k is needed in the cells within reach of a src flag.  A flag of level m in cell c' makes level k
// necessary in cells [c' - lo[m][k], c' + hi[m][k]] (both axes); the host derives the tables by interval arithmetic over
// the exact taps (make_reach_table / make_weight_reach_table).  One CTA per frame: the frame's src flags of all levels
// are packed into row bit masks in shared memory, a (level, row, 32-cell word) of the result is then a handful of
// shifts and ORs (a 2-D dilation), and its set bits are appended to the level's work list -- one global atomic per
// (CTA, level), items of a frame contiguous and in row-major order.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t shifted_word(const uint32_t* row, int w, int ww, int d) {   // bit x of the result = bit (32 w + x + d) of the row
    const uint32_t cur = row[w];
    if (d == 0) return cur;
    if (d > 0) return (cur >> d) | ((w + 1 < ww ? row[w + 1] : 0u) << (32 - d));
    return (cur << (-d)) | ((w > 0 ? row[w - 1] : 0u) >> (32 + d));
}

__global__ void __launch_bounds__(256) mbx_propagate_kernel(const __grid_constant__ GroupParams p, int image) {
    extern __shared__ uint32_t sbits[];   // [L][ch][ww]
    __shared__ unsigned s_count[6], s_base[6];
    const int f = blockIdx.x;
    const FrameJob& J = p.jobs[f];
    const int cw = J.wnx * 8, ch = J.wny * 8, ww = (cw + 31) >> 5, L = p.levels;
    const uint8_t* src = image ? p.win : p.comp;
    if (threadIdx.x < 6) s_count[threadIdx.x] = 0;
    const int nwords = L * ch * ww;
    for (int i = threadIdx.x; i < nwords; i += 256) {
        const int w = i % ww, y = (i / ww) % ch, m = i / (ww * ch);
        const uint8_t* q = src + cell_base(p, f, m) + (size_t)y * cw + w * 32;   // cell_base and cw are multiples of 8
        const int nb = min(32, cw - w * 32);
        uint32_t bits = 0u;
        for (int b8 = 0; b8 < nb; b8 += 8) {
            const uint2 v = *reinterpret_cast<const uint2*>(q + b8);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                bits |= (((v.x >> (8 * j)) & 0xFFu) ? 1u : 0u) << (b8 + j);
                bits |= (((v.y >> (8 * j)) & 0xFFu) ? 1u : 0u) << (b8 + 4 + j);
            }
        }
        sbits[i] = bits;
    }
    __syncthreads();
    // pass A: result words + their slot inside the CTA's share of the list
    constexpr int kMaxPerThread = 8;      // ceil(6 levels * 256 rows * 8 words / 256 threads) would be 48; regions that large take more rounds
    for (int i0 = 0; i0 < nwords; i0 += 256 * kMaxPerThread) {
        uint32_t acc[kMaxPerThread];
        unsigned off[kMaxPerThread];
#pragma unroll
        for (int r = 0; r < kMaxPerThread; r++) {
            const int i = i0 + r * 256 + threadIdx.x;
            acc[r] = 0u; off[r] = 0u;
            if (i >= nwords) continue;
            const int w = i % ww, y = (i / ww) % ch, k = i / (ww * ch);
            uint32_t a = 0u;
            for (int m = 0; m < L; m++) {
                const int rl = image ? p.reach_lo[m][k] : p.wreach_lo[m][k], rh = image ? p.reach_hi[m][k] : p.wreach_hi[m][k];
                if (rl == 0xFF) continue;
                // cell c is required by a flag in c' iff c' - rl <= c <= c' + rh  <=>  c - rh <= c' <= c + rl
                const int y0 = max(y - rh, 0), y1 = min(y + rl, ch - 1);
                for (int yy = y0; yy <= y1; yy++) {
                    const uint32_t* row = sbits + ((size_t)m * ch + yy) * ww;
                    for (int d = -rh; d <= rl; d++) a |= shifted_word(row, w, ww, d);
                }
            }
            const int nb = min(32, cw - w * 32);
            if (nb < 32) a &= (1u << nb) - 1u;
            acc[r] = a;
            if (a) off[r] = atomicAdd(&s_count[k], (unsigned)__popc(a));
        }
        __syncthreads();
        if (threadIdx.x < L && s_count[threadIdx.x]) {
            s_base[threadIdx.x] = atomicAdd(p.list_count + image * 6 + threadIdx.x, s_count[threadIdx.x]);
            if (p.need_stats) {   // px of level k the frames have to compute (a cell is (32 >> k)^2 px)
                const unsigned long long side = (unsigned long long)max(32 >> threadIdx.x, 1);
                atomicAdd(p.need_stats + (image ? 20 : 26) + threadIdx.x, side * side * (unsigned long long)s_count[threadIdx.x]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kMaxPerThread; r++) {
            uint32_t a = acc[r];
            if (!a) continue;
            const int i = i0 + r * 256 + threadIdx.x;
            const int w = i % ww, y = (i / ww) % ch, k = i / (ww * ch);
            uint32_t* dst = p.lists + ((size_t)image * L + k) * p.list_cap + s_base[k] + off[r];
            const uint32_t item0 = ((uint32_t)f << 16) | (uint32_t)(y * cw + w * 32);
            while (a) {
                const int b = __ffs(a) - 1;
                a &= a - 1;
                *dst++ = item0 + (uint32_t)b;
            }
        }
        __syncthreads();
        if (threadIdx.x < 6) s_count[threadIdx.x] = 0;
        __syncthreads();
    }
}
cudaError_t launch_mbx_propagate(const GroupParams& p, int image, cudaStream_t stream) {
    const int ww = (p.max_wnx * 8 + 31) / 32;
    const size_t smem = (size_t)p.levels * (p.max_wny * 8) * ww * sizeof(uint32_t);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;   // a frame region beyond ~250 x 250 tiles
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(mbx_propagate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    mbx_propagate_kernel<<<p.n_frames, 256, smem, stream>>>(p, image);
    return cudaGetLastError();
}

// Reach table of the image side.  A win of level m occupying cell c (level-m px [c*B, (c+1)*B - 1], B = 32 >> m)
// needs: its own px of G_m; if m is not the top level, G_{m+1} on [(lo >> 1) - 1, (hi >> 1) + 1] (the pyrUp taps of
// lap_quad); and every G_{k+1} px u needs G_k on [2u - 2, 2u + 2] (pyrDown taps; borders reflect inwards only).
// Cell of level-k px q = (q << k) >> 5.  The table is translation invariant (cells are aligned at every level).
void make_reach_table(int levels, unsigned char lo_tab[6][6], unsigned char hi_tab[6][6]) {
    for (int m = 0; m < 6; m++)
        for (int k = 0; k < 6; k++) lo_tab[m][k] = hi_tab[m][k] = 0xFF;
    const long long c = 1 << 12;
    for (int m = 0; m < levels && m < 6; m++) {
        const long long B = 32 >> m;
        long long a = c * B, b = (c + 1) * B - 1;
        int k = m;
        auto put = [&](int lvl, long long x0, long long x1) {
            long long c0 = (x0 << lvl) >> 5, c1 = (x1 << lvl) >> 5;
            lo_tab[m][lvl] = (unsigned char)(c - c0);
            hi_tab[m][lvl] = (unsigned char)(c1 - c);
        };
        if (m + 1 < levels) {
            a = (a >> 1) - 1; b = (b >> 1) + 1;
            k = m + 1;
        }
        put(k, a, b);
        for (; k > 0; k--) {
            a = 2 * a - 2; b = 2 * b + 2;
            put(k - 1, a, b);
        }
    }
}
// Reach table of the weight side: a competitive cell c of level m needs W_m on its own px and, down the pyrDown chain,
// W_{k} on [2u - 2, 2u + 2] for every needed px u of W_{k+1}.
void make_weight_reach_table(int levels, unsigned char lo_tab[6][6], unsigned char hi_tab[6][6]) {
    for (int m = 0; m < 6; m++)
        for (int k = 0; k < 6; k++) lo_tab[m][k] = hi_tab[m][k] = 0xFF;
    const long long c = 1 << 12;
    for (int m = 0; m < levels && m < 6; m++) {
        const long long B = 32 >> m;
        long long a = c * B, b = (c + 1) * B - 1;
        for (int k = m; k >= 0; k--) {
            long long c0 = (a << k) >> 5, c1 = (b << k) >> 5;
            lo_tab[m][k] = (unsigned char)(c - c0);
            hi_tab[m][k] = (unsigned char)(c1 - c);
            a = 2 * a - 2; b = 2 * b + 2;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// persistent per-frame stages over the work lists: item = (frame << 16) | cell of the frame's window
// ---------------------------------------------------------------------------------------------------------
// 1a. level-0 weights (nearest, constant 0; FP32 coordinate with exact FP64 redo of ambiguous px, see mbw_weights4):
// one CTA iteration = one 32 x 32 cell, one thread = 4 px
__global__ void __launch_bounds__(256) mbw_warp_kernel(const __grid_constant__ GroupParams p) {
    const unsigned n = p.list_count[0];
    const uint32_t* list = p.lists;
    for (unsigned it = blockIdx.x; it < n; it += gridDim.x) {
        const uint32_t item = list[it];
        const int f = (int)(item >> 16), c = (int)(item & 0xFFFFu);
        const FrameJob& J = p.jobs[f];
        const int cw = J.wnx * 8, cy = c / cw, cx = c - cy * cw, ww = J.wnx * kEle;
        const int u = cx * 32 + (threadIdx.x & 7) * 4, v = cy * 32 + (threadIdx.x >> 3);
        const float4 w = mbw_weights4(p, J, u + J.wx * kEle, v + J.wy * kEle);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.scratch + J.w_off[0]) + (size_t)v * ww + u) = w;
    }
}
cudaError_t launch_mbw_warp(const GroupParams& p, int ctas, cudaStream_t stream) {
    mbw_warp_kernel<<<ctas, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// 4a. level-0 Gaussian (u8x4, 1/32-px bilinear, BORDER_REFLECT) of the listed cells
__global__ void __launch_bounds__(256) mbs_warp_kernel(const __grid_constant__ GroupParams p) {
    const unsigned n = p.list_count[6];
    const uint32_t* list = p.lists + (size_t)p.levels * p.list_cap;
    for (unsigned it = blockIdx.x; it < n; it += gridDim.x) {
        const uint32_t item = list[it];
        const int f = (int)(item >> 16), c = (int)(item & 0xFFFFu);
        const FrameJob& J = p.jobs[f];
        const int cw = J.wnx * 8, cy = c / cw, cx = c - cy * cw, ww = J.wnx * kEle;
        const int u = cx * 32 + (threadIdx.x & 7) * 4, v = cy * 32 + (threadIdx.x >> 3);
        double M[9];
#pragma unroll
        for (int i = 0; i < 9; i++) M[i] = J.hinv[i];
        const RawSrc R = make_raw_src(J.raw, J.raw_stride, nullptr, p.sw, p.sh);
        uint32_t g[4];
        mb_sample4<false>(p, R, M, u + J.wx * kEle, v + J.wy * kEle, g, nullptr);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(p.scratch + J.g_off[0]) + (size_t)v * ww + u) = make_uint4(g[0], g[1], g[2], g[3]);
    }
}
cudaError_t launch_mbs_warp(const GroupParams& p, int ctas, cudaStream_t stream) {
    mbs_warp_kernel<<<ctas, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---- pull mode: host frames in pinned memory, weights-first pipeline ----
// The winners are known before a single image px has been touched, and a frame's image is needed only in the listed level-0
// cells.  So instead of staging whole frames (cudaMemcpyAsync: 2.76 MB per 720p frame, PCIe-bound: the whole e2e time), the GPU
// PULLS what it needs: mbs_mark maps every listed cell through the frame's inverse homography and marks the 256-byte chunks of
// the (row-major, tightly packed) frame that its bilinear taps can touch; mbs_pull copies the marked chunks from the caller's
// pinned host memory (device-visible under UVA) into the frame's staging slot in HBM with 16-byte lanes, 256 contiguous bytes
// per half-warp -- every needed byte crosses PCIe exactly once, in full-size read requests -- and mbs_warp then samples the
// slot as if the whole frame were there.  Conservative by construction (bounds.h pull_cell_rect: -2 / +3 px around the projected
// cell, reflection folded in, whole frame for a non-positive denominator; proven on the CPU against a brute-force walk of the
// taps), +-16 bytes for the aligned 3-word tap fetches.
__global__ void __launch_bounds__(256) mbs_mark_kernel(const __grid_constant__ GroupParams p) {
    const unsigned n = p.list_count[6];
    const uint32_t* list = p.lists + (size_t)p.levels * p.list_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long fb = (long long)p.frame_bytes;
    for (unsigned it = blockIdx.x * 8 + warp; it < n; it += gridDim.x * 8) {
        const uint32_t item = list[it];
        const int f = (int)(item >> 16), c = (int)(item & 0xFFFFu);
        const FrameJob& J = p.jobs[f];
        const int cw = J.wnx * 8, cy = c / cw, cx = c - cy * cw;
        int lox, hix, loy, hiy;
        pull_cell_rect(J.hinv, cx * 32 + J.wx * kEle, cy * 32 + J.wy * kEle, p.sw, p.sh, &lox, &hix, &loy, &hiy);   // bounds.h
        uint32_t* bits = p.src_bits + (size_t)f * p.src_words;
        for (int y = loy + lane; y <= hiy; y += 32) {
            long long b0 = (long long)y * J.raw_stride + 3 * lox - 16, b1 = (long long)y * J.raw_stride + 3 * hix + 2 + 16;
            b0 = b0 < 0 ? 0 : b0; b1 = b1 >= fb ? fb - 1 : b1;
            for (long long ch = b0 >> 8; ch <= (b1 >> 8); ch++) {
                uint32_t* wp = bits + (ch >> 5);
                const uint32_t m = 1u << (ch & 31);
                if (!(*reinterpret_cast<volatile uint32_t*>(wp) & m)) atomicOr(wp, m);
            }
        }
    }
}
cudaError_t launch_mbs_mark(const GroupParams& p, int ctas, cudaStream_t stream) {
    mbs_mark_kernel<<<ctas, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) mbs_pull_kernel(const __grid_constant__ GroupParams p) {
    const long long total = (long long)p.n_frames * p.src_words;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * 8;
    const size_t fb = (size_t)p.frame_bytes;
    for (long long idx = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); idx < total; idx += nwarps) {
        uint32_t bits = p.src_bits[idx];
        if (!bits) continue;
        const int f = (int)(idx / p.src_words), wi = (int)(idx - (long long)f * p.src_words);
        const FrameJob& J = p.jobs[f];
        const uint8_t* __restrict__ src = J.pull_src;
        uint8_t* __restrict__ dst = const_cast<uint8_t*>(J.raw);
        while (bits) {   // two chunks per pass: lanes 0-15 take the first set bit, lanes 16-31 the second
            const int b0 = __ffs(bits) - 1;
            bits &= bits - 1;
            int b1 = -1;
            if (bits) { b1 = __ffs(bits) - 1; bits &= bits - 1; }
            const int b = lane < 16 ? b0 : b1;
            if (b < 0) continue;
            const size_t off = (((size_t)wi * 32 + b) << 8) + (size_t)(lane & 15) * 16;
            if (off + 16 <= fb) *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(src + off);
            else for (size_t k = off; k < fb; k++) dst[k] = src[k];
        }
    }
}
cudaError_t launch_mbs_pull(const GroupParams& p, int ctas, cudaStream_t stream) {
    mbs_pull_kernel<<<ctas, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// 1b./4b. pyrDown l -> l+1 of the listed cells of level l+1 (a cell is S x S outputs, S = 32 >> (l+1)).
// Separable [1 4 6 4 1], BORDER_REFLECT_101 in REGION coordinates.  One thread = 2 adjacent output columns x up to 4
// output rows: the input rows stream through a 5-row register window, each fetched with 8-byte loads (7 input columns).
// Image: packed 16-bit lanes (B,R in one register, G alone): row sums <= 4080, column sums <= 65280, (x+128)>>8.
// Weight: f32 in OpenCV 2.4.9's association (rows s0*6 + (s-1 + s1)*4 + s-2 + s2 left to right; columns
// ((r0+r4)+(r2+r2)) + ((r1+r3)+r2)*4, scaled by 1/256 -- PyrDownVec_32f).
template <bool IMG>
__global__ void __launch_bounds__(256) mbx_pyrdown_kernel(const __grid_constant__ GroupParams p, int l) {
    const int S = 32 >> (l + 1);                     // 16, 8, 4, 2, 1
    const int cpw = max(S >> 1, 1), rqs = max(S >> 2, 1), tpc = cpw * rqs;   // threads per cell: 32, 8, 2, 1, 1
    const int ipc = 256 / tpc;                       // cells per CTA iteration
    const int rows = min(S, 4), rlim = 2 * rows + 3; // output rows per thread, input rows it streams
    const unsigned n = p.list_count[(IMG ? 6 : 0) + l + 1];
    const uint32_t* list = p.lists + ((size_t)(IMG ? p.levels : 0) + l + 1) * p.list_cap;
    const int sub = threadIdx.x / tpc, tt = threadIdx.x - sub * tpc;
    const int cp = tt % cpw, rq = tt / cpw;
    const int ns = kEle >> l, nd = kEle >> (l + 1);
    for (unsigned base = blockIdx.x * ipc; base < n; base += gridDim.x * ipc) {
        const unsigned it = base + sub;
        if (it >= n) continue;
        const uint32_t item = list[it];
        const int f = (int)(item >> 16), c = (int)(item & 0xFFFFu);
        const FrameJob& J = p.jobs[f];
        const int cw = J.wnx * 8, cy = c / cw, cx = c - cy * cw;
        const int sww = J.wnx * ns, swh = J.wny * ns, srw = J.nx * ns, srh = J.ny * ns, sox = J.wx * ns, soy = J.wy * ns;
        const int dww = J.wnx * nd, dox = J.wx * nd, doy = J.wy * nd;
        const int u = cx * S + 2 * cp, v0 = cy * S + 4 * rq;
        const bool two = S >= 2;
        const int U = u + dox, V0 = v0 + doy;
        int xs[7];
        const int c0 = 2 * U - 2 - sox;
        const bool fast = (2 * U - 2 >= 0) && (2 * U + 4 < srw) && (c0 >= 0) && (c0 + 6 < sww);
        if (fast) {
#pragma unroll
            for (int d = 0; d < 7; d++) xs[d] = c0 + d;
        } else {
#pragma unroll
            for (int d = 0; d < 7; d++) xs[d] = clampi(reflect101_idx(2 * U + d - 2, srw) - sox, 0, sww - 1);
        }
        if constexpr (IMG) {
            const uint32_t* SG = reinterpret_cast<const uint32_t*>(p.scratch + J.g_off[l]);
            uint32_t* DG = reinterpret_cast<uint32_t*>(p.scratch + J.g_off[l + 1]);
            uint32_t hb0[5], hg0[5], hb1[5], hg1[5];
#pragma unroll
            for (int r = 0; r < 11; r++) {
                if (r < rlim) {
                    const int ys = clampi(reflect101_idx(2 * V0 + r - 2, srh) - soy, 0, swh - 1);
                    const uint32_t* gr = SG + (size_t)ys * sww;
                    uint32_t e[7];
                    if (fast) {
                        const uint2* g2 = reinterpret_cast<const uint2*>(gr + xs[0]);
                        const uint2 a = g2[0], b = g2[1], cc = g2[2];
                        e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[4] = cc.x; e[5] = cc.y; e[6] = gr[xs[0] + 6];
                    } else {
#pragma unroll
                        for (int d = 0; d < 7; d++) e[d] = gr[xs[d]];
                    }
                    uint32_t br[7], g[7];
#pragma unroll
                    for (int d = 0; d < 7; d++) { br[d] = e[d] & kM2; g[d] = (e[d] >> 8) & 0xFFu; }
                    hb0[r % 5] = br[2] * 6u + (br[1] + br[3]) * 4u + br[0] + br[4];
                    hb1[r % 5] = br[4] * 6u + (br[3] + br[5]) * 4u + br[2] + br[6];
                    hg0[r % 5] = g[2] * 6u + (g[1] + g[3]) * 4u + g[0] + g[4];
                    hg1[r % 5] = g[4] * 6u + (g[3] + g[5]) * 4u + g[2] + g[6];
                    if (r >= 4 && (r & 1) == 0) {
                        const int k = (r - 4) >> 1, v = v0 + k;
                        const int i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
                        const uint32_t vbr0 = hb0[i0] + hb0[i4] + (hb0[i1] + hb0[i3]) * 4u + hb0[i2] * 6u;
                        const uint32_t vbr1 = hb1[i0] + hb1[i4] + (hb1[i1] + hb1[i3]) * 4u + hb1[i2] * 6u;
                        const uint32_t vg0 = hg0[i0] + hg0[i4] + (hg0[i1] + hg0[i3]) * 4u + hg0[i2] * 6u;
                        const uint32_t vg1 = hg1[i0] + hg1[i4] + (hg1[i1] + hg1[i3]) * 4u + hg1[i2] * 6u;
                        const uint32_t o0 = (((vbr0 + 0x00800080u) >> 8) & kM2) | (((vg0 + 128u) >> 8) << 8);
                        const uint32_t o1 = (((vbr1 + 0x00800080u) >> 8) & kM2) | (((vg1 + 128u) >> 8) << 8);
                        const size_t o = (size_t)v * dww + u;
                        if (two) *reinterpret_cast<uint2*>(DG + o) = make_uint2(o0, o1);   // u and dww are even when S >= 2
                        else DG[o] = o0;
                    }
                }
            }
        } else {
            const float* SW = reinterpret_cast<const float*>(p.scratch + J.w_off[l]);
            float* DW = reinterpret_cast<float*>(p.scratch + J.w_off[l + 1]);
            const F32Assoc assoc = f32_assoc(p.f32_mode, srw);
            float h0[5], h1[5];
#pragma unroll
            for (int r = 0; r < 11; r++) {
                if (r < rlim) {
                    const int ys = clampi(reflect101_idx(2 * V0 + r - 2, srh) - soy, 0, swh - 1);
                    const float* wr = SW + (size_t)ys * sww;
                    float fv[7];
                    if (fast) {
                        const float2* w2 = reinterpret_cast<const float2*>(wr + xs[0]);
                        const float2 fa = w2[0], fb = w2[1], fc = w2[2];
                        fv[0] = fa.x; fv[1] = fa.y; fv[2] = fb.x; fv[3] = fb.y; fv[4] = fc.x; fv[5] = fc.y; fv[6] = wr[xs[0] + 6];
                    } else {
#pragma unroll
                        for (int d = 0; d < 7; d++) fv[d] = wr[xs[d]];
                    }
                    h0[r % 5] = pyr_h(assoc, U, fv[0], fv[1], fv[2], fv[3], fv[4]);
                    h1[r % 5] = pyr_h(assoc, U + 1, fv[2], fv[3], fv[4], fv[5], fv[6]);
                    if (r >= 4 && (r & 1) == 0) {
                        const int k = (r - 4) >> 1, v = v0 + k;
                        const int i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
                        const float ow0 = pyr_v(assoc, U, h0[i0], h0[i1], h0[i2], h0[i3], h0[i4]), ow1 = pyr_v(assoc, U + 1, h1[i0], h1[i1], h1[i2], h1[i3], h1[i4]);
                        const size_t o = (size_t)v * dww + u;
                        if (two) *reinterpret_cast<float2*>(DW + o) = make_float2(ow0, ow1);
                        else DW[o] = ow0;
                    }
                }
            }
        }
    }
}
// ---- bulk-copy (TMA engine) helpers: 1-D cp.async.bulk global -> shared, completion counted on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {   // 16-byte aligned src, dst, size
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1b'./4b'. pyrDown 0 -> 1 of the listed cells with the input patch STAGED IN SHARED MEMORY BY THE TMA ENGINE.  The
// register-window kernel above is latency-bound (ncu: 31 % issue slots, 25-37 % DRAM, 33-49 % occupancy at 64-76
// registers): every thread waits on its own 44 global loads.  Here a warp owns a cell (16 x 16 outputs <- 35 x 35 input
// px): lane r issues ONE bulk copy per patch row (cp.async.bulk, 160 B = 40 px from the 16-byte aligned column
// 32 cx - 4) into the warp's stage buffer, completion is counted on the stage's mbarrier (expect_tx = 35 x 160 B), and
// the patch of the warp's NEXT cell is in flight while the current one is filtered out of shared memory (2 stages per
// warp, no CTA-wide barrier anywhere).  Same arithmetic, same association, same thread -> output mapping as the
// register-window kernel.  Cells on the rim of the frame's window (reflected or clamped taps, ~10 %) take the
// register-window path inside the same kernel.
constexpr int kPW = 40, kPH = 35, kPStages = 2;
constexpr int kPatchWords = kPW * kPH;                                   // 1400 words = 5600 B per stage
constexpr size_t kPyrTmaSmem = (size_t)8 * kPStages * kPatchWords * 4 + 8 * kPStages * 8;

template <bool IMG>
__global__ void __launch_bounds__(256) mbx_pyrdown0_tma_kernel(const __grid_constant__ GroupParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * kPStages * kPatchWords;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)8 * kPStages * kPatchWords * 4) + warp * kPStages;
    if (lane == 0) {
        for (int s = 0; s < kPStages; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const unsigned n = p.list_count[(IMG ? 6 : 0) + 1];
    const uint32_t* list = p.lists + ((size_t)(IMG ? p.levels : 0) + 1) * p.list_cap;
    const unsigned gw = blockIdx.x * 8 + warp, stride = gridDim.x * 8;
    const int cp = lane & 7, rq = lane >> 3;
    const int ns = kEle, nd = kEle >> 1;
    unsigned parity = 0;   // bit s: the phase parity the next wait on stage s expects

    // issue the patch of item `it` into stage s (all lanes call; rim cells issue nothing and return false)
    auto prefetch = [&](unsigned it, int s) -> bool {
        const uint32_t item = list[it];
        const int f = (int)(item >> 16), c = (int)(item & 0xFFFFu);
        const FrameJob& J = p.jobs[f];
        const int cw = J.wnx * 8, ch = J.wny * 8, cy = c / cw, cx = c - cy * cw;
        if (cx < 1 || cx > cw - 2 || cy < 1 || cy > ch - 2) return false;
        const int sww = J.wnx * ns;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.scratch + (IMG ? J.g_off[0] : J.w_off[0])) + (size_t)(32 * cy - 2) * sww + (32 * cx - 4);
        uint32_t* dst = buf + (size_t)s * kPatchWords;
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was read through the generic proxy two items ago
            mbar_arrive_expect_tx(&bars[s], kPH * kPW * 4);
        }
        __syncwarp();
        bulk_g2s(dst + lane * kPW, src + (size_t)lane * sww, kPW * 4, &bars[s]);
        if (lane < kPH - 32) bulk_g2s(dst + (32 + lane) * kPW, src + (size_t)(32 + lane) * sww, kPW * 4, &bars[s]);
        return true;
    };

    unsigned it = gw;
    bool staged = it < n ? prefetch(it, 0) : false;
    int s = 0;
    for (; it < n; it += stride, s ^= 1) {
        const unsigned nxt = it + stride;
        const bool staged_next = nxt < n ? prefetch(nxt, s ^ 1) : false;
        const uint32_t item = list[it];
        const int f = (int)(item >> 16), c = (int)(item & 0xFFFFu);
        const FrameJob& J = p.jobs[f];
        const int cw = J.wnx * 8, cy = c / cw, cx = c - cy * cw;
        const int dww = J.wnx * nd;
        const int u = cx * 16 + 2 * cp, v0 = cy * 16 + 4 * rq;
        const int UR = u + J.wx * nd;                              // output column in region coordinates
        const F32Assoc assoc = f32_assoc(p.f32_mode, J.nx * ns);
        if (staged) {
            mbar_wait(&bars[s], (parity >> s) & 1u);
            parity ^= 1u << s;
            const uint32_t* patch = buf + (size_t)s * kPatchWords + (size_t)(8 * rq) * kPW + 4 * cp + 2;   // the thread's 7 x 11 window
            if constexpr (IMG) {
                uint32_t* DG = reinterpret_cast<uint32_t*>(p.scratch + J.g_off[1]);
                uint32_t hb0[5], hg0[5], hb1[5], hg1[5];
#pragma unroll
                for (int r = 0; r < 11; r++) {
                    const uint32_t* gr = patch + r * kPW;
                    const uint2 a = *reinterpret_cast<const uint2*>(gr), b = *reinterpret_cast<const uint2*>(gr + 2), cc = *reinterpret_cast<const uint2*>(gr + 4);
                    const uint32_t e[7] = {a.x, a.y, b.x, b.y, cc.x, cc.y, gr[6]};
                    uint32_t br[7], g[7];
#pragma unroll
                    for (int d = 0; d < 7; d++) { br[d] = e[d] & kM2; g[d] = (e[d] >> 8) & 0xFFu; }
                    hb0[r % 5] = br[2] * 6u + (br[1] + br[3]) * 4u + br[0] + br[4];
                    hb1[r % 5] = br[4] * 6u + (br[3] + br[5]) * 4u + br[2] + br[6];
                    hg0[r % 5] = g[2] * 6u + (g[1] + g[3]) * 4u + g[0] + g[4];
                    hg1[r % 5] = g[4] * 6u + (g[3] + g[5]) * 4u + g[2] + g[6];
                    if (r >= 4 && (r & 1) == 0) {
                        const int k = (r - 4) >> 1, v = v0 + k;
                        const int i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
                        const uint32_t vbr0 = hb0[i0] + hb0[i4] + (hb0[i1] + hb0[i3]) * 4u + hb0[i2] * 6u;
                        const uint32_t vbr1 = hb1[i0] + hb1[i4] + (hb1[i1] + hb1[i3]) * 4u + hb1[i2] * 6u;
                        const uint32_t vg0 = hg0[i0] + hg0[i4] + (hg0[i1] + hg0[i3]) * 4u + hg0[i2] * 6u;
                        const uint32_t vg1 = hg1[i0] + hg1[i4] + (hg1[i1] + hg1[i3]) * 4u + hg1[i2] * 6u;
                        const uint32_t o0 = (((vbr0 + 0x00800080u) >> 8) & kM2) | (((vg0 + 128u) >> 8) << 8);
                        const uint32_t o1 = (((vbr1 + 0x00800080u) >> 8) & kM2) | (((vg1 + 128u) >> 8) << 8);
                        *reinterpret_cast<uint2*>(DG + (size_t)v * dww + u) = make_uint2(o0, o1);
                    }
                }
            } else {
                float* DW = reinterpret_cast<float*>(p.scratch + J.w_off[1]);
                const float* fpatch = reinterpret_cast<const float*>(patch);
                const bool odd = (rq & 1) != 0;
                float h0[5], h1[5];
#pragma unroll
                for (int r = 0; r < 11; r++) {
                    const float* wr = fpatch + r * kPW;
                    // The four row groups (rq) of a warp sit 8 * 40 words = a multiple of 32 banks apart, and every lane's first
                    // pair starts at a word = 2 (mod 4): loaded in the same order, the groups would collide 4-way.  Odd groups
                    // swap the order of each two pairs, so that at every LDS half of the lanes is on words 2,3 (mod 4) and the
                    // other half on words 0,1 -- 2-way instead of 4-way conflicts.  (The 4th pair's second word is never used.)
                    const float2 l0 = *reinterpret_cast<const float2*>(wr + (odd ? 2 : 0)), l1 = *reinterpret_cast<const float2*>(wr + (odd ? 0 : 2));
                    const float2 l2 = *reinterpret_cast<const float2*>(wr + (odd ? 6 : 4)), l3 = *reinterpret_cast<const float2*>(wr + (odd ? 4 : 6));
                    const float2 fa = odd ? l1 : l0, fb = odd ? l0 : l1, fc = odd ? l3 : l2;
                    const float fv[7] = {fa.x, fa.y, fb.x, fb.y, fc.x, fc.y, odd ? l2.x : l3.x};
                    h0[r % 5] = pyr_h(assoc, UR, fv[0], fv[1], fv[2], fv[3], fv[4]);
                    h1[r % 5] = pyr_h(assoc, UR + 1, fv[2], fv[3], fv[4], fv[5], fv[6]);
                    if (r >= 4 && (r & 1) == 0) {
                        const int k = (r - 4) >> 1, v = v0 + k;
                        const int i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
                        *reinterpret_cast<float2*>(DW + (size_t)v * dww + u) =
                            make_float2(pyr_v(assoc, UR, h0[i0], h0[i1], h0[i2], h0[i3], h0[i4]), pyr_v(assoc, UR + 1, h1[i0], h1[i1], h1[i2], h1[i3], h1[i4]));
                    }
                }
            }
            __syncwarp();   // every lane is done with the stage before it is refilled
        } else {
            // rim cell: reflected / clamped taps straight from global memory (register-window path)
            const int sww = J.wnx * ns, swh = J.wny * ns, srw = J.nx * ns, srh = J.ny * ns, sox = J.wx * ns, soy = J.wy * ns;
            const int U = u + J.wx * nd, V0 = v0 + J.wy * nd;
            int xs[7];
#pragma unroll
            for (int d = 0; d < 7; d++) xs[d] = clampi(reflect101_idx(2 * U + d - 2, srw) - sox, 0, sww - 1);
            if constexpr (IMG) {
                const uint32_t* SG = reinterpret_cast<const uint32_t*>(p.scratch + J.g_off[0]);
                uint32_t* DG = reinterpret_cast<uint32_t*>(p.scratch + J.g_off[1]);
                uint32_t hb0[5], hg0[5], hb1[5], hg1[5];
#pragma unroll
                for (int r = 0; r < 11; r++) {
                    const int ys = clampi(reflect101_idx(2 * V0 + r - 2, srh) - soy, 0, swh - 1);
                    const uint32_t* gr = SG + (size_t)ys * sww;
                    uint32_t br[7], g[7];
#pragma unroll
                    for (int d = 0; d < 7; d++) { const uint32_t e = gr[xs[d]]; br[d] = e & kM2; g[d] = (e >> 8) & 0xFFu; }
                    hb0[r % 5] = br[2] * 6u + (br[1] + br[3]) * 4u + br[0] + br[4];
                    hb1[r % 5] = br[4] * 6u + (br[3] + br[5]) * 4u + br[2] + br[6];
                    hg0[r % 5] = g[2] * 6u + (g[1] + g[3]) * 4u + g[0] + g[4];
                    hg1[r % 5] = g[4] * 6u + (g[3] + g[5]) * 4u + g[2] + g[6];
                    if (r >= 4 && (r & 1) == 0) {
                        const int k = (r - 4) >> 1, v = v0 + k;
                        const int i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
                        const uint32_t vbr0 = hb0[i0] + hb0[i4] + (hb0[i1] + hb0[i3]) * 4u + hb0[i2] * 6u;
                        const uint32_t vbr1 = hb1[i0] + hb1[i4] + (hb1[i1] + hb1[i3]) * 4u + hb1[i2] * 6u;
                        const uint32_t vg0 = hg0[i0] + hg0[i4] + (hg0[i1] + hg0[i3]) * 4u + hg0[i2] * 6u;
                        const uint32_t vg1 = hg1[i0] + hg1[i4] + (hg1[i1] + hg1[i3]) * 4u + hg1[i2] * 6u;
                        const uint32_t o0 = (((vbr0 + 0x00800080u) >> 8) & kM2) | (((vg0 + 128u) >> 8) << 8);
                        const uint32_t o1 = (((vbr1 + 0x00800080u) >> 8) & kM2) | (((vg1 + 128u) >> 8) << 8);
                        *reinterpret_cast<uint2*>(DG + (size_t)v * dww + u) = make_uint2(o0, o1);
                    }
                }
            } else {
                const float* SW = reinterpret_cast<const float*>(p.scratch + J.w_off[0]);
                float* DW = reinterpret_cast<float*>(p.scratch + J.w_off[1]);
                float h0[5], h1[5];
#pragma unroll
                for (int r = 0; r < 11; r++) {
                    const int ys = clampi(reflect101_idx(2 * V0 + r - 2, srh) - soy, 0, swh - 1);
                    const float* wr = SW + (size_t)ys * sww;
                    float fv[7];
#pragma unroll
                    for (int d = 0; d < 7; d++) fv[d] = wr[xs[d]];
                    h0[r % 5] = pyr_h(assoc, UR, fv[0], fv[1], fv[2], fv[3], fv[4]);
                    h1[r % 5] = pyr_h(assoc, UR + 1, fv[2], fv[3], fv[4], fv[5], fv[6]);
                    if (r >= 4 && (r & 1) == 0) {
                        const int k = (r - 4) >> 1, v = v0 + k;
                        const int i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
                        *reinterpret_cast<float2*>(DW + (size_t)v * dww + u) =
                            make_float2(pyr_v(assoc, UR, h0[i0], h0[i1], h0[i2], h0[i3], h0[i4]), pyr_v(assoc, UR + 1, h1[i0], h1[i1], h1[i2], h1[i3], h1[i4]));
                    }
                }
            }
        }
        staged = staged_next;
    }
}

cudaError_t launch_mbx_pyrdown(const GroupParams& p, int image, int level, int ctas, cudaStream_t stream) {
    if (level == 0 && p.use_tma) {
        static bool attr_done[64] = {};   // the attribute is per device (a multi-device handle launches on several)
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            cudaError_t e = cudaFuncSetAttribute(mbx_pyrdown0_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPyrTmaSmem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mbx_pyrdown0_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPyrTmaSmem);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
        const int tma_ctas = max(1, ctas / 4);   // 2 resident CTAs per SM (90 KB of stage buffers each)
        if (image) mbx_pyrdown0_tma_kernel<true><<<tma_ctas, 256, kPyrTmaSmem, stream>>>(p);
        else mbx_pyrdown0_tma_kernel<false><<<tma_ctas, 256, kPyrTmaSmem, stream>>>(p);
        return cudaGetLastError();
    }
    if (image) mbx_pyrdown_kernel<true><<<ctas, 256, 0, stream>>>(p, level);
    else mbx_pyrdown_kernel<false><<<ctas, 256, 0, stream>>>(p, level);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// 1c. entry table: for every (tile entry, level) the address of the frame's weight plane at the tile's origin and its row
// pitch, so that the decide stage reaches a candidate's weights with ONE dependent load (mask -> table -> weights) instead
// of three (mask -> entry -> frame job -> weights)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mbc_entry_table_kernel(const __grid_constant__ GroupParams p, int n_entries) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_entries * p.levels) return;
    const int e = i / p.levels, l = i - e * p.levels, n = kEle >> l;
    const TileEntry E = p.entries[e];
    const FrameJob& J = p.jobs[E.frame];
    const int st = J.wnx * n;
    EntryRef r;
    r.base = reinterpret_cast<unsigned long long>(p.scratch + J.w_off[l]) + 4ull * ((size_t)((E.rty - J.wy) * n) * st + (size_t)((E.rtx - J.wx) * n));
    r.stride = st;
    r.cell0 = (int)(((size_t)E.frame * p.levels + l) * p.cells_max + (size_t)((E.rty - J.wy) * 8) * (J.wnx * 8) + (E.rtx - J.wx) * 8);   // flag index of the tile's cell (0,0)
    p.etable[i] = r;
}
cudaError_t launch_mbc_entry_table(const GroupParams& p, int n_entries, cudaStream_t stream) {
    if (n_entries == 0) return cudaSuccess;
    mbc_entry_table_kernel<<<(n_entries * p.levels + 255) / 256, 256, 0, stream>>>(p, n_entries);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// 2. decide, tile-centric and cell-aligned: per px of the tile, scan the COMPETITIVE entries of its cell in feed order
// and keep the last one whose weight is >= everything before it (state included; a fresh tile starts at -inf so the
// first entry copies unconditionally, MultiBandMap2DCPU.cpp:498-504) -- exactly what the sequential `if (srcW >= dstW)`
// updates leave behind (:539-547; 0 >= 0 ties overwrite).  Winners go to the tile's weight plane, the winner map
// (entry index per pyramid px, 0xFFFF = the state stands) and the frames' `win` cell flags; cmin gets the exact new
// minimum of every revisited cell.  Cells without a competitive entry are skipped without a load.
// ---------------------------------------------------------------------------------------------------------
constexpr int kDecideChunk = 3;
__global__ void __launch_bounds__(256, 5) mbs_decide_kernel(const __grid_constant__ GroupParams p, const __grid_constant__ TileLayout lay) {
    const TileWork T = p.tiles[blockIdx.x];
    const CellMap m = cell_map(blockIdx.y, threadIdx.x, p.levels);
    if (!m.valid) return;
    const int l = m.l, n = m.n, px = m.px, py = m.py;
    const bool quad = m.npx == 4;
    const uint32_t* cm = p.cmask + (((size_t)blockIdx.x * p.levels + l) * 64 + m.cell) * p.mask_words;
    uint32_t any = 0u;
    for (int w = 0; w < p.mask_words; w++) any |= cm[w];
    if (!any) return;
    const size_t to = (size_t)py * n + px;
    float* tw = reinterpret_cast<float*>(T.state + lay.wgt_off[l]) + to;
    float bw[4];
    if (T.fresh) { bw[0] = bw[1] = bw[2] = bw[3] = -INFINITY; }
    else if (quad) {
        const float2 t0 = *reinterpret_cast<const float2*>(tw), t1 = *reinterpret_cast<const float2*>(tw + n);
        bw[0] = t0.x; bw[1] = t0.y; bw[2] = t1.x; bw[3] = t1.y;
    } else { bw[0] = tw[0]; bw[1] = bw[2] = bw[3] = INFINITY; }
    int best[4] = {-1, -1, -1, -1};
    unsigned wins = 0;
    // Candidates are taken kDecideChunk at a time: first their table entries, then their weights, only then the comparisons (in feed
    // order).  A cell rarely has more than three or four competitive frames, so all loads of a thread are in flight together instead
    // of one dependent chain (mask -> table -> weights) per candidate.
    for (int w = 0; w < p.mask_words; w++) {
        uint32_t bits = cm[w];
        while (bits) {
            int idx[kDecideChunk];
#pragma unroll
            for (int j = 0; j < kDecideChunk; j++) {
                idx[j] = -1;
                if (bits) { idx[j] = w * 32 + __ffs(bits) - 1; bits &= bits - 1; }
            }
            const float* qp[kDecideChunk];
            int st[kDecideChunk];
#pragma unroll
            for (int j = 0; j < kDecideChunk; j++) {
                qp[j] = nullptr; st[j] = 0;
                if (idx[j] >= 0) {
                    const EntryRef R = p.etable[(size_t)(T.first + idx[j]) * p.levels + l];
                    st[j] = R.stride;
                    qp[j] = reinterpret_cast<const float*>(R.base) + (size_t)py * R.stride + px;
                }
            }
            float2 t0[kDecideChunk], t1[kDecideChunk];
#pragma unroll
            for (int j = 0; j < kDecideChunk; j++) {
                t0[j] = make_float2(-INFINITY, -INFINITY); t1[j] = t0[j];
                if (idx[j] >= 0) {
                    if (quad) { t0[j] = *reinterpret_cast<const float2*>(qp[j]); t1[j] = *reinterpret_cast<const float2*>(qp[j] + st[j]); }
                    else t0[j].x = qp[j][0];
                }
            }
#pragma unroll
            for (int j = 0; j < kDecideChunk; j++) {
                if (idx[j] < 0) continue;
                const int i = idx[j];
                const unsigned cw = !(T.fresh && i == 0);
                if (quad) {
                    const float s[4] = {t0[j].x, t0[j].y, t1[j].x, t1[j].y};
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (s[k] >= bw[k]) { bw[k] = s[k]; best[k] = i; wins += cw; }   // '>=' : MultiBandMap2DCPU.cpp:542
                } else {
                    const float s = t0[j].x;
                    if (s >= bw[0]) { bw[0] = s; best[0] = i; wins += cw; }
                }
            }
        }
    }
    if (p.stats && wins) atomicAdd(p.stats + l, (unsigned long long)wins);
    // winner map + tile weights
    uint16_t* wm = p.wmap + (size_t)blockIdx.x * p.wmap_stride + lay.px_off[l] + to;
    if (!quad) {
        wm[0] = (uint16_t)(best[0] < 0 ? 0xFFFF : best[0]);
        if (best[0] >= 0) tw[0] = bw[0];
    } else {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int k0 = 2 * r, k1 = 2 * r + 1;
            *reinterpret_cast<ushort2*>(wm + (size_t)r * n) =
                make_ushort2((unsigned short)(best[k0] < 0 ? 0xFFFF : best[k0]), (unsigned short)(best[k1] < 0 ? 0xFFFF : best[k1]));
            float* wr = tw + (size_t)r * n;
            if (best[k0] >= 0 && best[k1] >= 0) *reinterpret_cast<float2*>(wr) = make_float2(bw[k0], bw[k1]);
            else if (best[k0] >= 0) wr[0] = bw[k0];
            else if (best[k1] >= 0) wr[1] = bw[k1];
        }
    }
    // exact minimum of the revisited cell (weights are >= 0, so the int order of the bit patterns is the float order)
    const float mn = fminf(fminf(bw[0], bw[1]), fminf(bw[2], bw[3]));
    atomicMin(reinterpret_cast<int*>(T.state + lay.cmin_off) + l * 64 + m.cell, __float_as_int(fmaxf(mn, 0.f)));
    // `win` cell flags of the winning frames (all px of a thread lie in one cell of the tile)
    const int ccx = m.cell & 7, ccy = m.cell >> 3;
    int seen[4] = {-1, -1, -1, -1};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (best[k] < 0) continue;
        bool dup = false;
#pragma unroll
        for (int k2 = 0; k2 < k; k2++) dup |= seen[k2] == best[k];
        seen[k] = best[k];
        if (dup) continue;
        const EntryRef R = p.etable[(size_t)(T.first + best[k]) * p.levels + l];
        uint8_t* wf = p.win + (size_t)(unsigned)R.cell0 + (size_t)ccy * (R.stride / n * 8) + ccx;   // stride / n = window width in tiles
        if (!*wf) *wf = 1;
    }
}
cudaError_t launch_mbs_decide(const GroupParams& p, const TileLayout& lay, cudaStream_t stream) {
    if (p.n_tiles == 0) return cudaSuccess;
    dim3 g(p.n_tiles, cellmap_ctas(p.levels));
    mbs_decide_kernel<<<g, 256, 0, stream>>>(p, lay);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// 5. Laplacian of the winners (read from the winner map) into the tile state; same mapping as decide, cells without a
// competitive entry (no winner map written) are skipped
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) mbs_lap_kernel(const __grid_constant__ GroupParams p, const __grid_constant__ TileLayout lay) {
    const TileWork T = p.tiles[blockIdx.x];
    const CellMap m = cell_map(blockIdx.y, threadIdx.x, p.levels);
    if (!m.valid) return;
    const int l = m.l, n = m.n, px = m.px, py = m.py;
    const bool quad = m.npx == 4;
    const uint32_t* cm = p.cmask + (((size_t)blockIdx.x * p.levels + l) * 64 + m.cell) * p.mask_words;
    uint32_t any = 0u;
    for (int w = 0; w < p.mask_words; w++) any |= cm[w];
    if (!any) return;
    const size_t to = (size_t)py * n + px;
    const uint16_t* wm = p.wmap + (size_t)blockIdx.x * p.wmap_stride + lay.px_off[l] + to;
    int best[4] = {-1, -1, -1, -1};
    if (quad) {
        const ushort2 a = *reinterpret_cast<const ushort2*>(wm), b = *reinterpret_cast<const ushort2*>(wm + n);
        best[0] = a.x == 0xFFFF ? -1 : a.x; best[1] = a.y == 0xFFFF ? -1 : a.y;
        best[2] = b.x == 0xFFFF ? -1 : b.x; best[3] = b.y == 0xFFFF ? -1 : b.y;
    } else best[0] = wm[0] == 0xFFFF ? -1 : wm[0];
    if (best[0] < 0 && best[1] < 0 && best[2] < 0 && best[3] < 0) return;
    int lap[4][3];
    bool done[4] = {best[0] < 0, best[1] < 0, best[2] < 0, best[3] < 0};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (done[k]) continue;
        const int f = best[k];
        const TileEntry E = p.entries[T.first + f];
        int tmp[4][3];
        lap_quad(p, p.jobs[E.frame], l, E.rtx * n + px, E.rty * n + py, quad, tmp);
#pragma unroll
        for (int k2 = k; k2 < 4; k2++)
            if (!done[k2] && best[k2] == f) { lap[k2][0] = tmp[k2][0]; lap[k2][1] = tmp[k2][1]; lap[k2][2] = tmp[k2][2]; done[k2] = true; }
    }
    const size_t plane = (size_t)n * n;
    int16_t* tl = reinterpret_cast<int16_t*>(T.state + lay.lap_off[l]) + to;
    if (!quad) {
        tl[0] = (int16_t)lap[0][0]; tl[plane] = (int16_t)lap[0][1]; tl[2 * plane] = (int16_t)lap[0][2];
        return;
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int k0 = 2 * r, k1 = 2 * r + 1;
        int16_t* tr = tl + (size_t)r * n;
        if (best[k0] >= 0 && best[k1] >= 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) *reinterpret_cast<short2*>(tr + c * plane) = make_short2((short)lap[k0][c], (short)lap[k1][c]);
        } else if (best[k0] >= 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) tr[c * plane] = (int16_t)lap[k0][c];
        } else if (best[k1] >= 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) tr[c * plane + 1] = (int16_t)lap[k1][c];
        }
    }
}
cudaError_t launch_mbs_lap(const GroupParams& p, const TileLayout& lay, cudaStream_t stream) {
    if (p.n_tiles == 0) return cudaSuccess;
    dim3 g(p.n_tiles, cellmap_ctas(p.levels));
    mbs_lap_kernel<<<g, 256, 0, stream>>>(p, lay);
    return cudaGetLastError();
}

}  // namespace m2d
