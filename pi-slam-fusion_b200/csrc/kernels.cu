// kernels.cu — hand-written sm_100a kernels of the Map2D feed() path.
//
// All arithmetic follows the exact integer / float recipes of the OpenCV primitives the reference calls
// (SURVEY.md §9): 1/32-px coordinate quantisation with round-half-even, fixed-point bilinear for 8UC4, exact
// bilinear + cvRound for 16SC3 with BORDER_REFLECT, nearest for the float weight, [1 4 6 4 1] pyrDown with
// (x+128)>>8, pyrUp with (x+32)>>6 and its asymmetric border.  Compiled with --fmad=false: no float or double
// contraction anywhere, so results match the CPU oracle bit for bit.
#include "kernels.cuh"

namespace m2d {

// ---------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int p, int len) {  // cv::borderInterpolate BORDER_REFLECT
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        p = (p < 0) ? (-p - 1) : (2 * len - 1 - p);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}
__device__ __forceinline__ int reflect101_idx(int p, int len) {  // BORDER_REFLECT_101
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        p = (p < 0) ? (-p) : (2 * len - 2 - p);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}
// BORDER_REFLECT for the common case of at most one fold per side, branch-free; falls back to the loop otherwise.
__device__ __forceinline__ int reflect_once(int p, int len) {
    int q = (p < 0) ? (-p - 1) : p;
    q = (q >= len) ? (2 * len - 1 - q) : q;
    if (__builtin_expect((unsigned)q >= (unsigned)len, 0)) q = reflect_idx(p, len);
    return q;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int sat16(int v) { return min(max(v, -32768), 32767); }

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// weight images — Map2DCPU.cpp:236-258 (u8 alpha) and MultiBandMap2DCPU.cpp:396-418 (f32)
// ---------------------------------------------------------------------------------------------------------
__global__ void weight_images_kernel(int sw, int sh, int weight_type, uint8_t* __restrict__ alpha,
                                     float* __restrict__ wimg) {
    int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= sw || i >= sh) return;
    float x_center = (float)(sw / 2), y_center = (float)(sh / 2);
    float dis_max = sqrtf(x_center * x_center + y_center * y_center);
    float dis = ((float)i - y_center) * ((float)i - y_center) + ((float)j - x_center) * ((float)j - x_center);
    dis = 1.f - sqrtf(dis) / dis_max;
    if (alpha) {
        int a;
        if (weight_type == 0) a = (int)((double)dis * 254.);
        else a = (int)(dis * dis * 254.f);
        a &= 255;
        if (a < 2) a = 2;
        alpha[(size_t)i * sw + j] = (uint8_t)a;
    }
    if (wimg) {
        float v = (weight_type == 0) ? dis : dis * dis;
        if ((double)v <= 1e-5) v = (float)1e-5;
        wimg[(size_t)i * sw + j] = v;
    }
}
cudaError_t launch_weight_images(int sw, int sh, int weight_type, uint8_t* alpha, float* wimg, cudaStream_t stream) {
    dim3 b(32, 8), g((sw + 31) / 32, (sh + 7) / 8);
    weight_images_kernel<<<g, b, 0, stream>>>(sw, sh, weight_type, alpha, wimg);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// bounds kernel — renderFrame part 1 for n poses (Map2DCPU.cpp:163-233 + getPerspectiveTransform + invert)
// ---------------------------------------------------------------------------------------------------------
__global__ void bounds_kernel(const __grid_constant__ GridGeom g, int n, const double* __restrict__ poses,
                              FrameBounds* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double pose[7];
    for (int i = 0; i < 7; i++) pose[i] = poses[(size_t)k * 7 + i];
    FrameBounds fb;
    frame_bounds(g, pose, &fb);
    out[k] = fb;
}
cudaError_t launch_bounds(const GridGeom& g, int n, const double* d_poses, FrameBounds* d_out, cudaStream_t stream) {
    bounds_kernel<<<(n + 63) / 64, 64, 0, stream>>>(g, n, d_poses, d_out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// shared sampling helpers
// ---------------------------------------------------------------------------------------------------------
// Row base of cv::warpPerspectiveInvoker for the 64-px block containing x: X0 = M0*xb + M1*y + M2, etc.
struct RowBase { double X0, Y0, W0; };
__device__ __forceinline__ RowBase row_base(const double* M, int x, int y) {
    int xb = x & ~63;
    RowBase r;
    r.X0 = M[0] * xb + M[1] * y + M[2];
    r.Y0 = M[3] * xb + M[4] * y + M[5];
    r.W0 = M[6] * xb + M[7] * y + M[8];
    return r;
}
// Un-quantised source coordinate of the px at offset x1 (as a double, exactly the int->double value OpenCV
// multiplies by) inside the block.  INTER_LINEAR rounds 32*f, INTER_NEAREST rounds f (32/W == 32*(1/W) exactly).
__device__ __forceinline__ void px_coord(const double* M, const RowBase& r, double x1, double& fx, double& fy) {
    double W = r.W0 + M[6] * x1;
    W = (W != 0.0) ? 1.0 / W : 0.0;
    fx = (r.X0 + M[0] * x1) * W;
    fy = (r.Y0 + M[3] * x1) * W;
}
// saturate_cast<int>(double): __double2int_rn rounds half to even and saturates, which equals OpenCV's
// max(INT_MIN, min(INT_MAX, f)) followed by cvRound for every non-NaN input.
__device__ __forceinline__ int rnd(double f) { return __double2int_rn(f); }
__device__ __forceinline__ int sat_s16(int v) { return min(max(v, -32768), 32767); }

constexpr uint32_t kM2 = 0x00FF00FFu;  // two 16-bit lanes holding one byte each

// ---------------------------------------------------------------------------------------------------------
// weighted mode, tile-centric: one CTA = 4 rows x 256 px of one tile, one thread = 4 consecutive px (one 16-byte
// state vector).  The thread walks the group's frames that touch the tile IN FEED ORDER, warps each (8UC4
// bilinear, constant-0 border, Map2DCPU.cpp:282-299) and keeps the strictly-larger alpha (Map2DCPU.cpp:324-329);
// the tile is read once and written once per group.  Frames are sampled in place (caller's BGR8 + alpha plane).
// ---------------------------------------------------------------------------------------------------------
// ---- weighted sampling straight from the caller's BGR8 frame + the alpha plane (no packed copy of the frame) ----
// One tap = 3 bytes at an arbitrary byte offset: fetch the aligned 32-bit words around it and funnel-shift.
struct RawSrc {
    const uint32_t* words;   // frame base rounded down to 4 bytes
    int mis;                 // base & 3
    int stride;              // bytes per row
    const uint8_t* alpha;    // sw*sh alpha plane (Map2DCPU.cpp:236-258)
    int sw, sh;
    unsigned last_word;      // index of the word holding the frame's last byte: the 3-word fetch never reads past it
};
__device__ __forceinline__ RawSrc make_raw_src(const uint8_t* raw, int stride, const uint8_t* alpha, int sw, int sh) {
    RawSrc R;
    R.mis = (int)(reinterpret_cast<uintptr_t>(raw) & 3);
    R.words = reinterpret_cast<const uint32_t*>(raw - R.mis);
    R.stride = stride; R.alpha = alpha; R.sw = sw; R.sh = sh;
    R.last_word = (unsigned)((sh - 1) * stride + 3 * sw - 1 + R.mis) >> 2;
    return R;
}
__device__ __forceinline__ uint32_t raw_tap(const RawSrc& R, int sx, int sy) {  // border-safe single tap (rare path)
    if ((unsigned)sx >= (unsigned)R.sw || (unsigned)sy >= (unsigned)R.sh) return 0u;
    const uint8_t* q = reinterpret_cast<const uint8_t*>(R.words) + R.mis + (size_t)sy * R.stride + 3 * sx;
    return (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16) |
           ((uint32_t)__ldg(R.alpha + sy * R.sw + sx) << 24);
}
// Two horizontally adjacent interior taps (sx, sx+1) of row sy: 6 consecutive bytes -> 3 aligned words.
__device__ __forceinline__ void raw_tap_pair(const RawSrc& R, int sx, int sy, uint32_t& v0, uint32_t& v1) {
    unsigned o = (unsigned)(sy * R.stride + 3 * sx + R.mis);
    const uint32_t* w = R.words + (o >> 2);
    unsigned sh = (o & 3u) * 8u;
    // the third word is only needed when bytes o+4/o+5 spill into it, and then it lies inside the frame: clamping
    // its index to the frame's last word therefore never changes a used byte, and never reads past the caller's buffer
    uint32_t lo = __ldg(w), mid = __ldg(w + 1), hi = __ldg(R.words + min((o >> 2) + 2u, R.last_word));
    uint32_t f0 = __funnelshift_r(lo, mid, sh), f1 = __funnelshift_r(mid, hi, sh);  // bytes o..o+3, o+4..o+7
    const uint8_t* ap = R.alpha + (sy * R.sw + sx);
    v0 = (f0 & 0x00FFFFFFu) | ((uint32_t)__ldg(ap) << 24);
    v1 = __byte_perm(f0, f1, 0x0543) & 0x00FFFFFFu;   // bytes o+3, o+4, o+5
    v1 |= (uint32_t)__ldg(ap + 1) << 24;
}

__device__ __forceinline__ uint32_t raw_tap_bgr(const RawSrc& R, int sx, int sy) {  // in-range single tap, no alpha
    const uint8_t* q = reinterpret_cast<const uint8_t*>(R.words) + R.mis + (size_t)sy * R.stride + 3 * sx;
    return (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
}
__device__ __forceinline__ void raw_tap_pair_bgr(const RawSrc& R, int sx, int sy, uint32_t& v0, uint32_t& v1) {
    unsigned o = (unsigned)(sy * R.stride + 3 * sx + R.mis);
    const uint32_t* w = R.words + (o >> 2);
    unsigned sh = (o & 3u) * 8u;
    uint32_t lo = __ldg(w), mid = __ldg(w + 1), hi = __ldg(R.words + min((o >> 2) + 2u, R.last_word));  // see raw_tap_pair
    uint32_t f0 = __funnelshift_r(lo, mid, sh), f1 = __funnelshift_r(mid, hi, sh);
    v0 = f0 & 0x00FFFFFFu;
    v1 = __byte_perm(f0, f1, 0x4543);   // bytes o+3, o+4, o+5, 0
}

// Returns the warped BGRA px, or 0 when its alpha cannot beat `cur_alpha` (colour math skipped).  `tie_wins`: an
// equal alpha also replaces (used when frames are visited out of feed order and this frame is the earlier one).
__device__ __forceinline__ uint32_t sample_bgra(const RawSrc& R, double fx, double fy, uint32_t cur_alpha, bool tie_wins, uint32_t& out_alpha) {
    const int sw = R.sw, sh = R.sh;
    int X = rnd(fx * 32.0), Y = rnd(fy * 32.0);
    int sx = X >> 5, sy = Y >> 5;
    if (__builtin_expect((unsigned)(X + 1048544) >= 2097088u || (unsigned)(Y + 1048544) >= 2097088u, 0)) {
        sx = sat_s16(sx); sy = sat_s16(sy);  // saturate_cast<short>: only beyond +-32767 px
    }
    out_alpha = 0;
    if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) return 0u;
    uint32_t v00, v01, v10, v11;
    if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
        raw_tap_pair(R, sx, sy, v00, v01);
        raw_tap_pair(R, sx, sy + 1, v10, v11);
    } else {
        v00 = raw_tap(R, sx, sy); v01 = raw_tap(R, sx + 1, sy);
        v10 = raw_tap(R, sx, sy + 1); v11 = raw_tap(R, sx + 1, sy + 1);
    }
    uint32_t a = X & 31, b = Y & 31, wa0 = 32 - a, wb0 = 32 - b;
    // horizontal pass on packed 16-bit lanes (max 255*32 = 8160 per lane); lanes (G,A) come out of one PRMT
    uint32_t ga0 = __byte_perm(v00, 0u, 0x4341) * wa0 + __byte_perm(v01, 0u, 0x4341) * a;
    uint32_t ga1 = __byte_perm(v10, 0u, 0x4341) * wa0 + __byte_perm(v11, 0u, 0x4341) * a;
    uint32_t A = ((ga0 >> 16) * wb0 + ((ga1 >> 16) * b + 512u)) >> 10;  // == (sum*32 + 16384) >> 15 (FixedPtCast<int,uchar,15>)
    out_alpha = A;
    if (A < cur_alpha || (A == cur_alpha && !tie_wins) || A == 0u) return 0u;
    uint32_t br0 = (v00 & kM2) * wa0 + (v01 & kM2) * a;
    uint32_t br1 = (v10 & kM2) * wa0 + (v11 & kM2) * a;
    uint32_t B = (__byte_perm(br0, 0u, 0x4410) * wb0 + (__byte_perm(br1, 0u, 0x4410) * b + 512u)) >> 10;
    uint32_t R_ = ((br0 >> 16) * wb0 + ((br1 >> 16) * b + 512u)) >> 10;
    uint32_t G = (__byte_perm(ga0, 0u, 0x4410) * wb0 + (__byte_perm(ga1, 0u, 0x4410) * b + 512u)) >> 10;
    return B | (G << 8) | (R_ << 16) | (A << 24);
}

// Upper bound of the warped alpha over a 4-px group whose source positions run from (ax,ay) to (bx,by): every tap
// lies within 1.5 px of that segment and the alpha image decreases with the distance to the frame centre
// (Map2DCPU.cpp:243-256), so alpha <= alpha(distance(centre, segment) - 1.5).  Conservative by construction (+1).
__device__ __forceinline__ uint32_t alpha_upper_bound(float ax, float ay, float bx, float by, float xc, float yc, float inv_dmax, int weight_type) {
    float vx = bx - ax, vy = by - ay, cx = xc - ax, cy = yc - ay;
    float vv = vx * vx + vy * vy;
    float t = vv > 0.f ? __fdividef(cx * vx + cy * vy, vv) : 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    float dx = cx - t * vx, dy = cy - t * vy;
    float r = fmaxf(sqrtf(dx * dx + dy * dy) - 1.5f, 0.f);
    float dis = fminf(1.f - r * inv_dmax + 1e-4f, 1.f);
    if (dis <= 0.f) return 2u;
    float v = weight_type == 0 ? dis * 254.f : dis * dis * 254.f;
    return (uint32_t)v + 2u;
}

// FP32 image of region px (x, y) under the inverse homography: only used for conservative culling (error << 1 px).
__device__ __forceinline__ void proj_f32(const float* __restrict__ m, float x, float y, float& sx, float& sy) {
    float w = m[6] * x + m[7] * y + m[8];
    float r = __fdividef(1.f, w);
    sx = (m[0] * x + m[1] * y + m[2]) * r;
    sy = (m[3] * x + m[4] * y + m[5]) * r;
}

__global__ void __launch_bounds__(256) weighted_group_kernel(const __grid_constant__ GroupParams p) {
    const TileWork T = p.tiles[blockIdx.x];
    const int lane = threadIdx.x & 31;
    int px = (threadIdx.x & 63) * 4, py = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int wpx0 = px & 128;  // a warp covers 128 consecutive px of one tile row
    uint4* sp = reinterpret_cast<uint4*>(T.state + ((size_t)py * kEle + px) * 4);
    uint4 st = T.fresh ? make_uint4(0u, 0u, 0u, 0u) : *sp;
    uint32_t s[4] = {st.x, st.y, st.z, st.w};
    // Frame (feed-order index) that currently holds each px; -1 = the state that was there before this group, which
    // wins every tie.  Needed because entries may be visited best-first instead of in feed order: the result of the
    // reference's sequential `if (tile.a < dst.a)` is "largest alpha, earliest frame on ties", which is order-free.
    int who[4] = {-1, -1, -1, -1};
    bool changed = T.fresh != 0;
    unsigned wins = 0, foot = 0;
    const float lim_x = (float)p.sw + 0.25f, lim_y = (float)p.sh + 0.25f;
    const float xc = (float)(p.sw / 2), yc = (float)(p.sh / 2);
    const float inv_dmax = rsqrtf(xc * xc + yc * yc) * 0.9999f;  // smaller => larger (conservative) alpha bound
    const bool cull_by_alpha = p.stats == nullptr;  // the counters follow the sequential semantics: no shortcuts then

    for (int c0 = 0; c0 < T.count; c0 += 32) {
        // ---- warp-level filter: lane i judges entry c0+i for the warp's whole 128-px row segment, in FP32 ----
        uint32_t ub = 0xFFFFFFFFu;   // alpha upper bound of "my" entry over the segment
        bool alive = c0 + lane < T.count;
        if (alive && cull_by_alpha) {
            const TileEntry E = p.entries[T.first + c0 + lane];
            const float* m = p.jobs[E.frame].hinvf;
            float X0 = (float)(E.rtx * kEle + wpx0), Y0 = (float)(E.rty * kEle + py), ax, ay, bx, by;
            proj_f32(m, X0, Y0, ax, ay);
            proj_f32(m, X0 + 127.f, Y0, bx, by);
            // margins: 1.5 px tap reach (see alpha_upper_bound) + 1 px for the FP32 coordinate error
            bool off = (ax < -2.25f && bx < -2.25f) || (ax > lim_x + 1.f && bx > lim_x + 1.f) || (ay < -2.25f && by < -2.25f) || (ay > lim_y + 1.f && by > lim_y + 1.f);
            // alpha_upper_bound already allows 1.5 px of tap reach; +2 alpha levels cover the FP32 coordinate error
            // (~0.01 px; the alpha image changes by 254/dmax per px)
            ub = alpha_upper_bound(ax, ay, bx, by, xc, yc, inv_dmax, p.weight_type) + 2u;
            alive = !off;
        }
        uint32_t amin = min(min(s[0] >> 24, s[1] >> 24), min(s[2] >> 24, s[3] >> 24));
        uint32_t wmin = cull_by_alpha ? __reduce_min_sync(0xffffffffu, amin) : 0u;
        unsigned mask = __ballot_sync(0xffffffffu, alive && ub >= wmin);
        while (mask) {
            const int i = __ffs(mask) - 1;
            mask &= mask - 1;
            const int e = c0 + i;
            // ---- exact path for entry e (per thread: 4 px) ----
            const TileEntry E = p.entries[T.first + e];
            const FrameJob& J = p.jobs[E.frame];
            int X = E.rtx * kEle + px, Y = E.rty * kEle + py;
            double M[9];
#pragma unroll
            for (int k = 0; k < 9; k++) M[k] = J.hinv[k];
            RowBase rb = row_base(M, X, Y);
            double x1 = (double)(X & 63);
            double fx[4], fy[4];
            px_coord(M, rb, x1, fx[0], fy[0]);
            px_coord(M, rb, x1 + 3.0, fx[3], fy[3]);
            // The 4 px lie on a line in the source too: if both ends are off the same side (with a margin far above
            // the rounding error) every tap of every px is outside the frame -> all four warp to 0.
            float ax = (float)fx[0], bx = (float)fx[3], ay = (float)fy[0], by = (float)fy[3];
            bool off = (ax < -1.25f && bx < -1.25f) || (ax > lim_x && bx > lim_x) || (ay < -1.25f && by < -1.25f) || (ay > lim_y && by > lim_y);
            bool skip = off;
            if (!skip && cull_by_alpha) {
                uint32_t a4 = min(min(s[0] >> 24, s[1] >> 24), min(s[2] >> 24, s[3] >> 24));
                skip = alpha_upper_bound(ax, ay, bx, by, xc, yc, inv_dmax, p.weight_type) < a4;  // cannot win or tie
            }
            if (!skip) {
                px_coord(M, rb, x1 + 1.0, fx[1], fy[1]);
                px_coord(M, rb, x1 + 2.0, fx[2], fy[2]);
                const RawSrc R = make_raw_src(J.raw, J.raw_stride, p.alpha, p.sw, p.sh);
                bool count_wins = !(T.fresh && e == 0);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t alpha;
                    bool tie_wins = who[j] >= 0 && E.frame < who[j];
                    uint32_t d = sample_bgra(R, fx[j], fy[j], s[j] >> 24, tie_wins, alpha);
                    foot += alpha != 0u;
                    if (d) {  // alpha beats the holder's (strict '<', Map2DCPU.cpp:327), or equals it and this frame is earlier
                        s[j] = d;
                        who[j] = E.frame;
                        changed = true;
                        wins += count_wins;
                    }
                }
            }
            // the state only improves: re-filter the entries still queued against the new segment minimum
            if (cull_by_alpha && mask) {
                amin = min(min(s[0] >> 24, s[1] >> 24), min(s[2] >> 24, s[3] >> 24));
                wmin = __reduce_min_sync(0xffffffffu, amin);
                mask &= __ballot_sync(0xffffffffu, ub >= wmin);
            }
        }
    }
    if (changed) *sp = make_uint4(s[0], s[1], s[2], s[3]);
    if (p.stats) {
        unsigned long long f = warp_sum(foot), w = warp_sum(wins);
        if (lane == 0) {
            if (f) atomicAdd(p.stats + 16, f);
            if (w) atomicAdd(p.stats + 17, w);
        }
    }
}
cudaError_t launch_weighted_group(const GroupParams& p, cudaStream_t stream) {
    if (p.n_tiles == 0) return cudaSuccess;
    dim3 g(p.n_tiles, kEle / 4);
    weighted_group_kernel<<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// multi-band stage 1: warp every frame of the group into level 0 of its scratch pyramid (over its window).
//   image : 16SC3 bilinear, BORDER_REFLECT, exact integer form of remapBilinear<Cast<float,short>> + cvRound; the
//           result is always in [0,255] so it is stored as packed u8x4 (B,G,R,0)
//   weight: nearest from the float weight image, constant-0 border
// grid = (256 px x 4 rows blocks, frame); one thread = 4 consecutive px.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bilinear_rne_bgr(uint32_t v00, uint32_t v01, uint32_t v10, uint32_t v11, uint32_t a, uint32_t b) {
    uint32_t wa0 = 32 - a, wb0 = 32 - b;
    // horizontal pass: B,R on packed 16-bit lanes (<= 8160), G alone; byte 1 extracted with one PRMT
    uint32_t br0 = (v00 & kM2) * wa0 + (v01 & kM2) * a, br1 = (v10 & kM2) * wa0 + (v11 & kM2) * a;
    uint32_t g0 = __byte_perm(v00, 0u, 0x4441) * wa0 + __byte_perm(v01, 0u, 0x4441) * a;
    uint32_t g1 = __byte_perm(v10, 0u, 0x4441) * wa0 + __byte_perm(v11, 0u, 0x4441) * a;
    // vertical pass with the +511 of the rounding folded into the multiply-add chain
    uint32_t B = __byte_perm(br0, 0u, 0x4410) * wb0 + (__byte_perm(br1, 0u, 0x4410) * b + 511u);
    uint32_t R = __byte_perm(br0, 0u, 0x4432) * wb0 + (__byte_perm(br1, 0u, 0x4432) * b + 511u);
    uint32_t G = g0 * wb0 + (g1 * b + 511u);
    // the float sum S00*w0+S01*w1+S10*w2+S11*w3 is exact (<= 18 bits), so cvRound(sum) == RNE(v / 1024):
    // (v + 511 + bit10(v)) >> 10, with bit10(v) = bit10((v+511) - 511)
    B = (B + (((B - 511u) >> 10) & 1u)) >> 10;
    G = (G + (((G - 511u) >> 10) & 1u)) >> 10;
    R = (R + (((R - 511u) >> 10) & 1u)) >> 10;
    return B | (G << 8) | (R << 16);
}

// Warp 4 consecutive region px (x..x+3 on row y; x is a multiple of 4, so they share one 64-px coordinate block).
template <bool WGT = true>
__device__ __forceinline__ void mb_sample4(const GroupParams& p, const RawSrc& R, const double* M, int x, int y, uint32_t* g, float* w) {
    RowBase rb = row_base(M, x, y);
    double x1 = (double)(x & 63);
    const int sw = R.sw, sh = R.sh;
    const float* __restrict__ wimg = p.wimg;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double fx, fy;
        px_coord(M, rb, x1 + (double)j, fx, fy);
        int X = rnd(fx * 32.0), Y = rnd(fy * 32.0);
        int nx = 0, ny = 0;
        if constexpr (WGT) { nx = rnd(fx); ny = rnd(fy); }
        // saturate_cast<short> of the integer coordinates only matters beyond +-32767 px: test once, clamp rarely
        if (__builtin_expect((unsigned)(X + 1048544) >= 2097088u || (unsigned)(Y + 1048544) >= 2097088u, 0)) {
            nx = sat_s16(nx); ny = sat_s16(ny);
            X = (sat_s16(X >> 5) << 5) | (X & 31); Y = (sat_s16(Y >> 5) << 5) | (Y & 31);
        }
        int sx = X >> 5, sy = Y >> 5;
        if constexpr (WGT) w[j] = ((unsigned)nx < (unsigned)sw && (unsigned)ny < (unsigned)sh) ? __ldg(wimg + (ny * sw + nx)) : 0.f;
        uint32_t v00, v01, v10, v11;
        if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
            raw_tap_pair_bgr(R, sx, sy, v00, v01);
            raw_tap_pair_bgr(R, sx, sy + 1, v10, v11);
        } else {
            int sx0 = reflect_once(sx, sw), sx1 = reflect_once(sx + 1, sw), sy0 = reflect_once(sy, sh), sy1 = reflect_once(sy + 1, sh);
            v00 = raw_tap_bgr(R, sx0, sy0); v01 = raw_tap_bgr(R, sx1, sy0); v10 = raw_tap_bgr(R, sx0, sy1); v11 = raw_tap_bgr(R, sx1, sy1);
        }
        // a = X & 31 as X - 32*sx: an IMAD on the FMA pipe instead of a LOP3 on the (saturated) ALU pipe
        g[j] = bilinear_rne_bgr(v00, v01, v10, v11, (uint32_t)(X - 32 * sx), (uint32_t)(Y - 32 * sy));
    }
}

__global__ void __launch_bounds__(256) mb_warp_kernel(const __grid_constant__ GroupParams p) {
    const FrameJob& J = p.jobs[blockIdx.y];
    const int ww = J.wnx * kEle, wh = J.wny * kEle;
    const int bpr = J.wnx;  // 256-px blocks per window row
    int by = blockIdx.x / bpr, bx = blockIdx.x - by * bpr;
    if (by * 4 >= wh) return;
    int u = bx * kEle + (threadIdx.x & 63) * 4, v = by * 4 + (threadIdx.x >> 6);
    int x = u + J.wx * kEle, y = v + J.wy * kEle;  // region coordinates
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; i++) M[i] = J.hinv[i];
    const RawSrc R = make_raw_src(J.raw, J.raw_stride, nullptr, p.sw, p.sh);
    uint32_t g[4];
    float w[4];
    mb_sample4(p, R, M, x, y, g, w);
    size_t o = (size_t)v * ww + u;
    *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(p.scratch + J.g_off[0]) + o) = make_uint4(g[0], g[1], g[2], g[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.scratch + J.w_off[0]) + o) = make_float4(w[0], w[1], w[2], w[3]);
}
cudaError_t launch_mb_warp(const GroupParams& p, cudaStream_t stream) {
    dim3 g(p.max_wnx * p.max_wny * (kEle / 4), p.n_frames);
    mb_warp_kernel<<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// multi-band stages 1+2a fused: warp a 128 x 32 px block of level 0 PLUS the halo the first pyrDown needs into shared
// memory, write the block to the level-0 scratch, and produce its 64 x 16 px of level 1 straight from shared memory.
// The level-0 scratch (12.6 MB per 720p frame) is then never read back by a pyrDown pass; the price is re-warping
// the halo (136 x 35 instead of 128 x 32 samples).  Borders: the pyrDown taps are reflected (BORDER_REFLECT_101) in
// REGION coordinates, which always lands inside the block itself, so px outside the region are never sampled.
// ---------------------------------------------------------------------------------------------------------
constexpr int kFW = 128, kFH = 32;            // level-0 block
constexpr int kFSW = kFW + 8, kFSH = kFH + 3;  // shared tile: columns x0-4 .. x0+131 (4-px aligned groups), rows y0-2 .. y0+32

__global__ void __launch_bounds__(256, 5) mb_warp_pyr_kernel(const __grid_constant__ GroupParams p) {
    __shared__ uint32_t sG[kFSH][kFSW];
    __shared__ float sW[kFSH][kFSW];
    const FrameJob& J = p.jobs[blockIdx.y];
    const int ww = J.wnx * kEle, wh = J.wny * kEle, rw = J.nx * kEle, rh = J.ny * kEle;
    const int bpr = ww / kFW;
    int by = blockIdx.x / bpr, bx = blockIdx.x - by * bpr;
    if (by * kFH >= wh) return;
    const int x0 = bx * kFW + J.wx * kEle, y0 = by * kFH + J.wy * kEle;  // block origin, region coordinates
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; i++) M[i] = J.hinv[i];
    const RawSrc R = make_raw_src(J.raw, J.raw_stride, nullptr, p.sw, p.sh);
    uint32_t* G0 = reinterpret_cast<uint32_t*>(p.scratch + J.g_off[0]);
    float* W0 = reinterpret_cast<float*>(p.scratch + J.w_off[0]);
    // ---- phase 1: sample the block + halo, 4-px groups
    constexpr int kGroups = (kFSW / 4) * kFSH;
    for (int gi = threadIdx.x; gi < kGroups; gi += 256) {
        int r = gi / (kFSW / 4), c = (gi - r * (kFSW / 4)) * 4;
        int x = x0 - 4 + c, y = y0 - 2 + r;
        if (x < 0 || x >= rw || y < 0 || y >= rh) continue;  // outside the region: reflected taps never read it
        uint32_t g[4];
        float w[4];
        mb_sample4(p, R, M, x, y, g, w);
        *reinterpret_cast<uint4*>(&sG[r][c]) = make_uint4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(&sW[r][c]) = make_float4(w[0], w[1], w[2], w[3]);
        if (c >= 4 && c < 4 + kFW && r >= 2 && r < 2 + kFH) {  // the block itself goes to the level-0 scratch
            size_t o = (size_t)(y - J.wy * kEle) * ww + (x - J.wx * kEle);
            *reinterpret_cast<uint4*>(G0 + o) = make_uint4(g[0], g[1], g[2], g[3]);
            *reinterpret_cast<float4*>(W0 + o) = make_float4(w[0], w[1], w[2], w[3]);
        }
    }
    __syncthreads();
    // ---- phase 2: level 1 of the block from shared memory: thread = 1 output column x 4 output rows
    const int dww = ww >> 1;
    uint32_t* G1 = reinterpret_cast<uint32_t*>(p.scratch + J.g_off[1]);
    float* W1 = reinterpret_cast<float*>(p.scratch + J.w_off[1]);
    const int ul = threadIdx.x & 63, vl0 = (threadIdx.x >> 6) * 4;
    const int U = (x0 >> 1) + ul, V0 = (y0 >> 1) + vl0;  // region coordinates at level 1
    int cs[5];
#pragma unroll
    for (int d = 0; d < 5; d++) cs[d] = reflect101_idx(2 * U + d - 2, rw) - (x0 - 4);
    uint32_t hbr[5], hg[5];  // 5-row sliding window of horizontal sums
    float hw[5];
#pragma unroll
    for (int r = 0; r < 11; r++) {
        int rr = reflect101_idx(2 * V0 + r - 2, rh) - (y0 - 2);
        const uint32_t* gr = sG[rr];
        const float* wr = sW[rr];
        uint32_t a = gr[cs[0]], b = gr[cs[1]], c = gr[cs[2]], d = gr[cs[3]], e = gr[cs[4]];
        hbr[r % 5] = (c & kM2) * 6u + ((b & kM2) + (d & kM2)) * 4u + (a & kM2) + (e & kM2);
        hg[r % 5] = ((c >> 8) & 0xFFu) * 6u + (((b >> 8) & 0xFFu) + ((d >> 8) & 0xFFu)) * 4u + ((a >> 8) & 0xFFu) + ((e >> 8) & 0xFFu);
        // f32, OpenCV 2.4.9 association: s0*6 + (s-1 + s1)*4 + s-2 + s2, left to right
        hw[r % 5] = wr[cs[2]] * 6.f + (wr[cs[1]] + wr[cs[3]]) * 4.f + wr[cs[0]] + wr[cs[4]];
        if (r >= 4 && (r & 1) == 0) {
            const int k = (r - 4) >> 1, i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
            uint32_t vbr = hbr[i0] + hbr[i4] + (hbr[i1] + hbr[i3]) * 4u + hbr[i2] * 6u;
            uint32_t vg = hg[i0] + hg[i4] + (hg[i1] + hg[i3]) * 4u + hg[i2] * 6u;
            // columns ((r0+r4)+(r2+r2)) + ((r1+r3)+r2)*4, scaled by 1/256 (PyrDownVec_32f of OpenCV 2.4.9)
            float t0 = (hw[i0] + hw[i4]) + (hw[i2] + hw[i2]);
            float t1 = (hw[i1] + hw[i3]) + hw[i2];
            size_t o = (size_t)(V0 + k - ((J.wy * kEle) >> 1)) * dww + (U - ((J.wx * kEle) >> 1));
            G1[o] = (((vbr + 0x00800080u) >> 8) & kM2) | (((vg + 128u) >> 8) << 8);
            W1[o] = (t0 + t1 * 4.f) * (1.f / 256.f);
        }
    }
}
cudaError_t launch_mb_warp_pyr(const GroupParams& p, cudaStream_t stream) {
    dim3 g(p.max_wnx * p.max_wny * (kEle / kFW) * (kEle / kFH), p.n_frames);
    mb_warp_pyr_kernel<<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// multi-band stage 2: pyrDown level l -> l+1 for every frame of the group (u8x4 Gaussian + f32 weight).
// Separable [1 4 6 4 1], BORDER_REFLECT_101 in REGION coordinates.  The int path runs on packed 16-bit lanes
// (B,R in one register, G alone): row sums <= 4080, column sums <= 65280, (x+128)>>8 -- no lane ever overflows.
// One thread = 2 adjacent output columns x 4 output rows: 11 input rows stream through a 5-row register window,
// each fetched with four 8-byte loads per plane (7 input columns), so every input px is loaded ~1.1x.
// ---------------------------------------------------------------------------------------------------------
struct HRow { uint32_t br0, g0, br1, g1; float w0, w1; };  // horizontal sums for output columns u and u+1

__device__ __forceinline__ HRow pyr_hrow(const uint32_t* __restrict__ gr, const float* __restrict__ wr, const int* xs, bool fast) {
    uint32_t e[7];
    float f[7];
    if (fast) {  // xs[0] is even and the 7 columns are consecutive: 8-byte vector loads
        const uint2* g2 = reinterpret_cast<const uint2*>(gr + xs[0]);
        const float2* w2 = reinterpret_cast<const float2*>(wr + xs[0]);
        uint2 a = g2[0], b = g2[1], c = g2[2];
        float2 fa = w2[0], fb = w2[1], fc = w2[2];
        e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[4] = c.x; e[5] = c.y; e[6] = gr[xs[0] + 6];
        f[0] = fa.x; f[1] = fa.y; f[2] = fb.x; f[3] = fb.y; f[4] = fc.x; f[5] = fc.y; f[6] = wr[xs[0] + 6];
    } else {
#pragma unroll
        for (int d = 0; d < 7; d++) { e[d] = gr[xs[d]]; f[d] = wr[xs[d]]; }
    }
    uint32_t br[7], g[7];
#pragma unroll
    for (int d = 0; d < 7; d++) { br[d] = e[d] & kM2; g[d] = (e[d] >> 8) & 0xFFu; }
    HRow h;
    h.br0 = br[2] * 6u + (br[1] + br[3]) * 4u + br[0] + br[4];
    h.br1 = br[4] * 6u + (br[3] + br[5]) * 4u + br[2] + br[6];
    h.g0 = g[2] * 6u + (g[1] + g[3]) * 4u + g[0] + g[4];
    h.g1 = g[4] * 6u + (g[3] + g[5]) * 4u + g[2] + g[6];
    // f32, OpenCV 2.4.9 association: s0*6 + (s-1 + s1)*4 + s-2 + s2, left to right
    h.w0 = f[2] * 6.f + (f[1] + f[3]) * 4.f + f[0] + f[4];
    h.w1 = f[4] * 6.f + (f[3] + f[5]) * 4.f + f[2] + f[6];
    return h;
}

__global__ void __launch_bounds__(256, 4) mb_pyrdown_kernel(const __grid_constant__ GroupParams p, int l) {
    const FrameJob& J = p.jobs[blockIdx.y];
    const int ns = kEle >> l, nd = kEle >> (l + 1);
    const int sww = J.wnx * ns, swh = J.wny * ns, srw = J.nx * ns, srh = J.ny * ns, sox = J.wx * ns, soy = J.wy * ns;
    const int dww = J.wnx * nd, dwh = J.wny * nd, dox = J.wx * nd, doy = J.wy * nd;
    const int bpr = (dww + 63) / 64;
    int by = blockIdx.x / bpr, bx = blockIdx.x - by * bpr;
    int u = (bx * 32 + threadIdx.x) * 2, v0 = (by * 8 + threadIdx.y) * 4;
    if (u >= dww || v0 >= dwh) return;
    const uint32_t* SG = reinterpret_cast<const uint32_t*>(p.scratch + J.g_off[l]);
    const float* SW = reinterpret_cast<const float*>(p.scratch + J.w_off[l]);
    uint32_t* DG = reinterpret_cast<uint32_t*>(p.scratch + J.g_off[l + 1]);
    float* DW = reinterpret_cast<float*>(p.scratch + J.w_off[l + 1]);
    const bool two = (u + 1) < dww;  // dww is odd only for the 1-px-wide top levels
    int U = u + dox;
    int xs[7];
    int c0 = 2 * U - 2 - sox;
    const bool fast = (2 * U - 2 >= 0) && (2 * U + 4 < srw) && (c0 >= 0) && (c0 + 6 < sww);
    if (fast) {
#pragma unroll
        for (int d = 0; d < 7; d++) xs[d] = c0 + d;
    } else {
#pragma unroll
        for (int d = 0; d < 7; d++) xs[d] = clampi(reflect101_idx(2 * U + d - 2, srw) - sox, 0, sww - 1);
    }
    int V0 = v0 + doy;
    HRow h[5];
#pragma unroll
    for (int r = 0; r < 11; r++) {
        int ys = clampi(reflect101_idx(2 * V0 + r - 2, srh) - soy, 0, swh - 1);
        h[r % 5] = pyr_hrow(SG + (size_t)ys * sww, SW + (size_t)ys * sww, xs, fast);
        if (r >= 4 && (r & 1) == 0) {
            int k = (r - 4) >> 1, v = v0 + k;
            if (v < dwh) {
                const HRow &r0 = h[(r - 4) % 5], &r1 = h[(r - 3) % 5], &r2 = h[(r - 2) % 5], &r3 = h[(r - 1) % 5], &r4 = h[r % 5];
                uint32_t vbr0 = r0.br0 + r4.br0 + (r1.br0 + r3.br0) * 4u + r2.br0 * 6u;
                uint32_t vbr1 = r0.br1 + r4.br1 + (r1.br1 + r3.br1) * 4u + r2.br1 * 6u;
                uint32_t vg0 = r0.g0 + r4.g0 + (r1.g0 + r3.g0) * 4u + r2.g0 * 6u;
                uint32_t vg1 = r0.g1 + r4.g1 + (r1.g1 + r3.g1) * 4u + r2.g1 * 6u;
                uint32_t o0 = (((vbr0 + 0x00800080u) >> 8) & kM2) | (((vg0 + 128u) >> 8) << 8);
                uint32_t o1 = (((vbr1 + 0x00800080u) >> 8) & kM2) | (((vg1 + 128u) >> 8) << 8);
                // columns ((r0+r4)+(r2+r2)) + ((r1+r3)+r2)*4, scaled by 1/256 (PyrDownVec_32f of OpenCV 2.4.9)
                float t00 = (r0.w0 + r4.w0) + (r2.w0 + r2.w0), t10 = (r1.w0 + r3.w0) + r2.w0;
                float t01 = (r0.w1 + r4.w1) + (r2.w1 + r2.w1), t11 = (r1.w1 + r3.w1) + r2.w1;
                float ow0 = (t00 + t10 * 4.f) * (1.f / 256.f), ow1 = (t01 + t11 * 4.f) * (1.f / 256.f);
                size_t o = (size_t)v * dww + u;
                if (two && !(dww & 1)) {  // 8-byte stores need an even row pitch (odd only at 1-px tile levels)
                    *reinterpret_cast<uint2*>(DG + o) = make_uint2(o0, o1);
                    *reinterpret_cast<float2*>(DW + o) = make_float2(ow0, ow1);
                } else {
                    DG[o] = o0;
                    DW[o] = ow0;
                    if (two) { DG[o + 1] = o1; DW[o + 1] = ow1; }
                }
            }
        }
    }
}
cudaError_t launch_mb_pyrdown(const GroupParams& p, int level, cudaStream_t stream) {
    // grid.x bound: the widest / tallest window of the group at level+1, in 64 x 32 output blocks
    int nd = kEle >> (level + 1);
    int blocks = ((p.max_wnx * nd + 63) / 64) * ((p.max_wny * nd + 31) / 32);
    dim3 b(32, 8), g(blocks, p.n_frames);
    mb_pyrdown_kernel<<<g, b, 0, stream>>>(p, level);
    return cudaGetLastError();
}

// Pyramid tail: the deepest levels are a few thousand px per frame -- too small for a launch each (a launch costs
// ~10 us of ramp/drain here).  One CTA per frame walks levels l_first..levels-2 in sequence, __syncthreads()
// between levels (the CTA itself wrote what it reads next).  Same arithmetic as mb_pyrdown_kernel, one px per thread.
__global__ void __launch_bounds__(1024) mb_pyrtail_kernel(const __grid_constant__ GroupParams p, int l_first) {
    const FrameJob& J = p.jobs[blockIdx.x];
    for (int l = l_first; l + 1 < p.levels; l++) {
        const int ns = kEle >> l, nd = kEle >> (l + 1);
        const int sww = J.wnx * ns, swh = J.wny * ns, srw = J.nx * ns, srh = J.ny * ns, sox = J.wx * ns, soy = J.wy * ns;
        const int dww = J.wnx * nd, dwh = J.wny * nd, dox = J.wx * nd, doy = J.wy * nd;
        const uint32_t* SG = reinterpret_cast<const uint32_t*>(p.scratch + J.g_off[l]);
        const float* SW = reinterpret_cast<const float*>(p.scratch + J.w_off[l]);
        uint32_t* DG = reinterpret_cast<uint32_t*>(p.scratch + J.g_off[l + 1]);
        float* DW = reinterpret_cast<float*>(p.scratch + J.w_off[l + 1]);
        for (int o = threadIdx.x; o < dww * dwh; o += blockDim.x) {
            int v = o / dww, u = o - v * dww;
            int U = u + dox, V = v + doy;
            int xs[5], ys[5];
#pragma unroll
            for (int d = 0; d < 5; d++) {
                xs[d] = clampi(reflect101_idx(2 * U + d - 2, srw) - sox, 0, sww - 1);
                ys[d] = clampi(reflect101_idx(2 * V + d - 2, srh) - soy, 0, swh - 1);
            }
            uint32_t hbr[5], hg[5];
            float hw[5];
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t* gr = SG + (size_t)ys[r] * sww;
                const float* wr = SW + (size_t)ys[r] * sww;
                uint32_t a = gr[xs[0]], b = gr[xs[1]], c = gr[xs[2]], d = gr[xs[3]], e = gr[xs[4]];
                hbr[r] = (c & kM2) * 6u + ((b & kM2) + (d & kM2)) * 4u + (a & kM2) + (e & kM2);
                hg[r] = ((c >> 8) & 0xFFu) * 6u + (((b >> 8) & 0xFFu) + ((d >> 8) & 0xFFu)) * 4u + ((a >> 8) & 0xFFu) + ((e >> 8) & 0xFFu);
                hw[r] = wr[xs[2]] * 6.f + (wr[xs[1]] + wr[xs[3]]) * 4.f + wr[xs[0]] + wr[xs[4]];
            }
            uint32_t vbr = hbr[0] + hbr[4] + (hbr[1] + hbr[3]) * 4u + hbr[2] * 6u;
            uint32_t vg = hg[0] + hg[4] + (hg[1] + hg[3]) * 4u + hg[2] * 6u;
            float t0 = (hw[0] + hw[4]) + (hw[2] + hw[2]);
            float t1 = (hw[1] + hw[3]) + hw[2];
            DG[o] = (((vbr + 0x00800080u) >> 8) & kM2) | (((vg + 128u) >> 8) << 8);
            DW[o] = (t0 + t1 * 4.f) * (1.f / 256.f);
        }
        __syncthreads();
    }
}
cudaError_t launch_mb_pyrtail(const GroupParams& p, int l_first, cudaStream_t stream) {
    mb_pyrtail_kernel<<<p.n_frames, 1024, 0, stream>>>(p, l_first);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// multi-band stage 3, tile-centric: per tile px (all levels in one launch) find the LAST frame of the group whose
// weight is >= everything before it (state included) -- exactly what sequential `if (srcW >= dstW)` updates leave
// behind (MultiBandMap2DCPU.cpp:539-547; 0 >= 0 ties overwrite) -- and only for that frame form the Laplacian
// G_l - pyrUp(G_{l+1}) (never materialised) and store it.  One thread = 2 horizontally adjacent px.
// ---------------------------------------------------------------------------------------------------------
TileLayout make_tile_layout(int levels) {
    TileLayout t{};
    t.levels = levels;
    size_t off = 0;
    int px = 0;
    for (int l = 0; l < levels; l++) {
        size_t n = (size_t)(kEle >> l);
        t.lap_off[l] = off;
        off += n * n * 3 * sizeof(int16_t);
        off = (off + 15) & ~(size_t)15;
        t.wgt_off[l] = off;
        off += n * n * sizeof(float);
        off = (off + 15) & ~(size_t)15;
        t.px_off[l] = px;
        px += (int)(n * n);
    }
    t.px_off[levels] = px;
    t.bytes = (off + 255) & ~(size_t)255;
    return t;
}

__device__ __forceinline__ int pyrup_axis_lo(int i, int n) { return i < 0 ? (n > 1 ? 1 : 0) : i; }  // reflect-101 at -1
__device__ __forceinline__ int pyrup_axis_hi(int i, int n) { return i >= n ? n - 1 : i; }          // replicate at n

// Laplacian G_l - pyrUp(G_{l+1}) of the 2x2 quad whose top-left px is (X, Y) (both even, region coordinates of
// level l) of frame J.  The four px share one 3x3 neighbourhood of the coarser level.  out[k][c]: k = 2*row + col.
__device__ __forceinline__ void lap_quad(const GroupParams& p, const FrameJob& J, int l, int X, int Y, bool quad, int out[4][3]) {
    const int n = kEle >> l;
    const int ww = J.wnx * n, ox = J.wx * n, oy = J.wy * n;
    const uint32_t* G = reinterpret_cast<const uint32_t*>(p.scratch + J.g_off[l]);
    size_t so = (size_t)(Y - oy) * ww + (X - ox);
    uint32_t g[4];
    if (quad) {
        uint2 r0 = *reinterpret_cast<const uint2*>(G + so), r1 = *reinterpret_cast<const uint2*>(G + so + ww);
        g[0] = r0.x; g[1] = r0.y; g[2] = r1.x; g[3] = r1.y;
    } else { g[0] = G[so]; g[1] = g[2] = g[3] = 0u; }
    if (l == p.levels - 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) { out[k][0] = g[k] & 0xFF; out[k][1] = (g[k] >> 8) & 0xFF; out[k][2] = (g[k] >> 16) & 0xFF; }
        return;
    }
    const int nc = n >> 1;
    const int cww = J.wnx * nc, cwh = J.wny * nc, crw = J.nx * nc, crh = J.ny * nc, cox = J.wx * nc, coy = J.wy * nc;
    const uint32_t* C = reinterpret_cast<const uint32_t*>(p.scratch + J.g_off[l + 1]);
    int i = X >> 1, j = Y >> 1;
    int c0 = clampi(pyrup_axis_lo(i - 1, crw) - cox, 0, cww - 1), c1 = clampi(i - cox, 0, cww - 1);
    int c2 = clampi(pyrup_axis_hi(i + 1, crw) - cox, 0, cww - 1);
    int r0 = clampi(pyrup_axis_lo(j - 1, crh) - coy, 0, cwh - 1), r1 = clampi(j - coy, 0, cwh - 1);
    int r2 = clampi(pyrup_axis_hi(j + 1, crh) - coy, 0, cwh - 1);
    const int rr[3] = {r0, r1, r2};
    uint32_t ebr[3], eg[3], obr[3], og[3];  // even / odd column sums per coarse row, packed lanes (<= 2040)
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint32_t* q = C + (size_t)rr[k] * cww;
        uint32_t a = q[c0], b = q[c1], c = q[c2];
        uint32_t abr = a & kM2, bbr = b & kM2, cbr = c & kM2, ag = (a >> 8) & 0xFFu, bg = (b >> 8) & 0xFFu, cg = (c >> 8) & 0xFFu;
        ebr[k] = abr + bbr * 6u + cbr; obr[k] = (bbr + cbr) * 4u;
        eg[k] = ag + bg * 6u + cg; og[k] = (bg + cg) * 4u;
    }
    // even output row: r0 + 6 r1 + r2 ; odd output row: 4 (r1 + r2)   (<= 16320 per lane), then (x + 32) >> 6
    uint32_t up_br[4], up_g[4];
    up_br[0] = (((ebr[0] + ebr[1] * 6u + ebr[2]) + 0x00200020u) >> 6) & kM2;
    up_br[1] = (((obr[0] + obr[1] * 6u + obr[2]) + 0x00200020u) >> 6) & kM2;
    up_br[2] = ((((ebr[1] + ebr[2]) * 4u) + 0x00200020u) >> 6) & kM2;
    up_br[3] = ((((obr[1] + obr[2]) * 4u) + 0x00200020u) >> 6) & kM2;
    up_g[0] = ((eg[0] + eg[1] * 6u + eg[2]) + 32u) >> 6;
    up_g[1] = ((og[0] + og[1] * 6u + og[2]) + 32u) >> 6;
    up_g[2] = (((eg[1] + eg[2]) * 4u) + 32u) >> 6;
    up_g[3] = (((og[1] + og[2]) * 4u) + 32u) >> 6;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        out[k][0] = (int)(g[k] & 0xFF) - (int)(up_br[k] & 0xFFFF);
        out[k][1] = (int)((g[k] >> 8) & 0xFF) - (int)up_g[k];
        out[k][2] = (int)((g[k] >> 16) & 0xFF) - (int)(up_br[k] >> 16);
    }
}

__global__ void __launch_bounds__(256, 6) mb_select_kernel(const __grid_constant__ GroupParams p, const __grid_constant__ TileLayout lay) {
    const TileWork T = p.tiles[blockIdx.x];
    // flat quad index -> (level, quad row, quad column); level l has (n/2)^2 quads (1 for the 1-px level)
    int q = blockIdx.y * 256 + threadIdx.x;
    int l = 0, n = kEle, half = kEle / 2;
    for (; l < p.levels; l++) {
        n = kEle >> l;
        half = n > 1 ? n / 2 : 1;
        int cnt = half * half;
        if (q < cnt) break;
        q -= cnt;
    }
    const bool valid = l < p.levels;
    if (!valid) { l = p.levels - 1; n = kEle >> l; half = n > 1 ? n / 2 : 1; q = 0; }
    const int qy = q / half, qx = q - qy * half;
    const int py = qy * 2, px = qx * 2;
    const bool quad = n > 1;
    const size_t to = (size_t)py * n + px;
    float* tw = reinterpret_cast<float*>(T.state + lay.wgt_off[l]) + to;
    float bw[4];
    if (T.fresh) { bw[0] = bw[1] = bw[2] = bw[3] = -INFINITY; }  // first toucher copies unconditionally (:498-504)
    else if (quad) {
        float2 t0 = *reinterpret_cast<const float2*>(tw), t1 = *reinterpret_cast<const float2*>(tw + n);
        bw[0] = t0.x; bw[1] = t0.y; bw[2] = t1.x; bw[3] = t1.y;
    } else { bw[0] = tw[0]; bw[1] = bw[2] = bw[3] = 0.f; }
    int best[4] = {-1, -1, -1, -1};
    unsigned wins = 0;
    const int lane = threadIdx.x & 31;
    // Levels 0..4 have a multiple of 32 quads, so a warp normally sits inside one level: then each lane resolves
    // ONE entry's weight-plane address (job lookup, window offset) and the warp shares them by shuffle.
    const bool uniform = __all_sync(0xffffffffu, l == __shfl_sync(0xffffffffu, l, 0));
    for (int c0 = 0; c0 < T.count; c0 += 32) {
        unsigned long long wb = 0ull;
        int stride = 0;
        if (uniform && c0 + lane < T.count) {
            const TileEntry E = p.entries[T.first + c0 + lane];
            const FrameJob& J = p.jobs[E.frame];
            stride = J.wnx * n;
            wb = reinterpret_cast<unsigned long long>(p.scratch + J.w_off[l]) +
                 4ull * ((size_t)((E.rty - J.wy) * n) * stride + (size_t)((E.rtx - J.wx) * n));
        }
        const int m = min(32, T.count - c0);
#pragma unroll 4
        for (int i = 0; i < m; i++) {
            const float* W;
            int st;
            if (uniform) {
                W = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, wb, i));
                st = __shfl_sync(0xffffffffu, stride, i);
            } else {
                const TileEntry E = p.entries[T.first + c0 + i];
                const FrameJob& J = p.jobs[E.frame];
                st = J.wnx * n;
                W = reinterpret_cast<const float*>(p.scratch + J.w_off[l]) + (size_t)((E.rty - J.wy) * n) * st + (size_t)((E.rtx - J.wx) * n);
            }
            const float* qp = W + (size_t)py * st + px;
            float s[4];
            if (quad) {
                float2 t0 = *reinterpret_cast<const float2*>(qp), t1 = *reinterpret_cast<const float2*>(qp + st);
                s[0] = t0.x; s[1] = t0.y; s[2] = t1.x; s[3] = t1.y;
            } else { s[0] = qp[0]; s[1] = s[2] = s[3] = -INFINITY; }
            const unsigned cw = !(T.fresh && (c0 + i) == 0);
#pragma unroll
            for (int k = 0; k < 4; k++)
                if ((quad || k == 0) && s[k] >= bw[k]) { bw[k] = s[k]; best[k] = c0 + i; wins += cw; }   // '>=' : MultiBandMap2DCPU.cpp:542
        }
    }
    if (!valid) { best[0] = best[1] = best[2] = best[3] = -1; wins = 0; }
    if (p.stats) {  // block-uniform branch: every lane reaches the shuffle
        unsigned long long w = warp_sum(wins);
        if (lane == 0 && w) atomicAdd(p.stats + l, w);
    }
    if ((best[0] & best[1] & best[2] & best[3]) < 0 && best[0] < 0 && best[1] < 0 && best[2] < 0 && best[3] < 0) return;

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
        tw[0] = bw[0];
        return;
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int k0 = 2 * r, k1 = 2 * r + 1;
        int16_t* tr = tl + (size_t)r * n;
        float* wr = tw + (size_t)r * n;
        if (best[k0] >= 0 && best[k1] >= 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) *reinterpret_cast<short2*>(tr + c * plane) = make_short2((short)lap[k0][c], (short)lap[k1][c]);
            *reinterpret_cast<float2*>(wr) = make_float2(bw[k0], bw[k1]);
        } else if (best[k0] >= 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) tr[c * plane] = (int16_t)lap[k0][c];
            wr[0] = bw[k0];
        } else if (best[k1] >= 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) tr[c * plane + 1] = (int16_t)lap[k1][c];
            wr[1] = bw[k1];
        }
    }
}
cudaError_t launch_mb_select(const GroupParams& p, const TileLayout& lay, cudaStream_t stream) {
    if (p.n_tiles == 0) return cudaSuccess;
    int quads = 0;
    for (int l = 0; l < p.levels; l++) {
        int n = kEle >> l, half = n > 1 ? n / 2 : 1;
        quads += half * half;
    }
    dim3 g(p.n_tiles, (quads + 255) / 256);
    mb_select_kernel<<<g, 256, 0, stream>>>(p, lay);
    return cudaGetLastError();
}

// =========================================================================================================
// multi-band, WEIGHTS-FIRST variant.  The weight pyramids depend on geometry only, so the winner of
// every pyramid px can be decided before a single image sample is taken; image warp and image pyrDown then run
// only where a winner's Laplacian needs them.  Stages per group:
//   1. mbw_warp / mbw_pyrdown / mbw_pyrtail   weights only, dense                      (order-independent)
//   2. mbs_decide                             tile-centric arg-max -> tile weights, winner map, `win` cell flags
//   3. mbs_propagate                          per frame: `need` = win, dilated down the dependency cone
//   4. mbs_warp / mbs_pyrdown / mbs_pyrtail   image only, only in needed cells
//   5. mbs_lap                                winner-only Laplacian -> tile state
// A CELL is 32 x 32 level-0 px of a frame's window; at level l it is (32 >> l)^2 px (requires levels <= 6), so one
// cell grid (8 x 8 cells per tile) serves every level: cell of level-l px p = (p << l) >> 5.  Dependencies in cell
// space: Laplacian l needs G_{l+1} at (p>>1)-1 .. (p>>1)+1 -> cells c-1..c+1; G_{l+1}(u) needs G_l at 2u-2 .. 2u+2
// -> cells c-1..c+1 (borders reflect inwards, never further).  Px of G outside needed cells are never read by a
// needed px, so they may hold anything.  Results are identical to the dense path.
// =========================================================================================================
__device__ __forceinline__ size_t cell_base(const GroupParams& p, int frame, int l) { return ((size_t)frame * p.levels + l) * p.cells_max; }

// ---- 1a. weight warp (nearest, constant 0): one thread = 4 px ----
// The weight of a px is wimg[rnd(fy)][rnd(fx)]: only the ROUNDED source coordinate matters.  So the coordinate is first
// evaluated in FP32; its distance to the exact FP64 value OpenCV computes is below a few ulps of the largest
// intermediate (bounded per thread by `mag`), hence rnd() of both agree unless the FP32 value lies within `thr` =
// 48 ulps(mag) of a rounding boundary (x.5).  Only those px (~2 %) take the exact FP64 path; lanes pick their own
// ambiguous px, so a warp normally runs that path once instead of four times.  Px far outside the frame skip both.
// Weights of 4 consecutive region px (x..x+3 on row y, x a multiple of 4) of frame J: FP32 pass + exact FP64 redo of the
// ambiguous px, as described above.
__device__ __forceinline__ float4 mbw_weights4(const GroupParams& p, const FrameJob& J, int x, int y) {
    float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
    const float* mf = J.hinvf;
    const float xf = (float)x, yf = (float)y;
    const float den0 = mf[7] * yf + mf[8], nx0 = mf[1] * yf + mf[2], ny0 = mf[4] * yf + mf[5];
    const float wa = mf[6] * xf + den0, wb = mf[6] * (xf + 3.f) + den0;
    const bool den_ok = wa > 1e-3f && wb > 1e-3f;   // denominators safely positive: FP32 reasoning is valid
    unsigned amb = 0xFu;                             // px that need the exact path
    if (den_ok) {
        const float lim_x = (float)p.sw + 0.25f, lim_y = (float)p.sh + 0.25f;
        float ra = __fdividef(1.f, wa), rb_ = __fdividef(1.f, wb);
        float ax = (mf[0] * xf + nx0) * ra, ay = (mf[3] * xf + ny0) * ra;
        float bx_ = (mf[0] * (xf + 3.f) + nx0) * rb_, by_ = (mf[3] * (xf + 3.f) + ny0) * rb_;
        // both ends of the 4-px run outside the same side of the source -> every px of the run is outside (a projective
        // map keeps the run a straight segment) -> weight 0
        const bool off = (ax < -1.25f && bx_ < -1.25f) || (ax > lim_x && bx_ > lim_x) || (ay < -1.25f && by_ < -1.25f) || (ay > lim_y && by_ > lim_y);
        if (off) amb = 0u;
        else {
            const float rmax = fmaxf(ra, rb_);
            const float magx = (fabsf(mf[0]) * (xf + 3.f) + fabsf(mf[1]) * yf + fabsf(mf[2])) * rmax;
            const float magy = (fabsf(mf[3]) * (xf + 3.f) + fabsf(mf[4]) * yf + fabsf(mf[5])) * rmax;
            const float thr_x = 48.f * 5.97e-8f * magx + 1e-6f, thr_y = 48.f * 5.97e-8f * magy + 1e-6f;
            amb = 0u;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float xj = xf + (float)j;
                const float r = __fdividef(1.f, mf[6] * xj + den0);
                const float fx = (mf[0] * xj + nx0) * r, fy = (mf[3] * xj + ny0) * r;
                const float rx = rintf(fx), ry = rintf(fy);
                const bool near_half = (0.5f - fabsf(fx - rx) < thr_x) || (0.5f - fabsf(fy - ry) < thr_y);
                float wv = 0.f;
                if (rx >= 0.f && rx < (float)p.sw && ry >= 0.f && ry < (float)p.sh) wv = __ldg(p.wimg + ((int)ry * p.sw + (int)rx));
                if (near_half) amb |= 1u << j;
                if (j == 0) w0 = wv; else if (j == 1) w1 = wv; else if (j == 2) w2 = wv; else w3 = wv;
            }
        }
    }
    if (amb) {   // exact OpenCV arithmetic for the px the FP32 pass could not decide
        double M[9];
#pragma unroll
        for (int i = 0; i < 9; i++) M[i] = J.hinv[i];
        RowBase rb = row_base(M, x, y);
        const double x1 = (double)(x & 63);
        while (amb) {
            const int j = __ffs(amb) - 1;
            amb &= amb - 1;
            double fx, fy;
            px_coord(M, rb, x1 + (double)j, fx, fy);
            int nx = rnd(fx), ny = rnd(fy);   // saturate_cast<short> cannot change an in/out decision for sw, sh <= 32767
            float wv = ((unsigned)nx < (unsigned)p.sw && (unsigned)ny < (unsigned)p.sh) ? __ldg(p.wimg + (ny * p.sw + nx)) : 0.f;
            if (j == 0) w0 = wv; else if (j == 1) w1 = wv; else if (j == 2) w2 = wv; else w3 = wv;
        }
    }
    return make_float4(w0, w1, w2, w3);
}

// EXPERIMENTAL (M2D_WLEAN=1, not the default, never run on a GPU yet): the same contract as mbw_weights4 with a shorter
// FP32 pass, aimed at what ncu shows for it (issue-bound; XU pipe 49 %, ALU 53 %, FP64 0.4 %):
//   * explicit FMAs (the file is compiled with --fmad=false, which otherwise splits every a*b+c);
//   * ONE reciprocal per run, carried to the next px by a Newton step r' = r + r*(1 - den'*r) on the FMA pipe (valid while
//     the denominator moves by < 3.3e-5 relative per px: the quadratic error term stays below 0.02 ulp);
//   * rounding by the 1.5*2^23 trick: t = f + 12582912 is f rounded half-to-even (|f| < 2^22), t - 12582912 the rounded
//     value and bits(t) - 0x4B400000 the integer, so no rintf / float->int conversion goes to the quarter-rate XU pipe;
//   * the off-frame test reuses the first and last px of the run.
// The ambiguity band (48 ulps of the largest intermediate) and the exact FP64 redo are unchanged; the numpy emulation in
// tests/test_weights_first_host.py checks that px the pass does not flag round like the exact FP64 path.
__device__ __forceinline__ float4 mbw_weights4_lean(const GroupParams& p, const FrameJob& J, int x, int y) {
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    const float* mf = J.hinvf;
    const float xf = (float)x, yf = (float)y;
    const float den0 = __fmaf_rn(mf[7], yf, mf[8]), nx0 = __fmaf_rn(mf[1], yf, mf[2]), ny0 = __fmaf_rn(mf[4], yf, mf[5]);
    const float wa = __fmaf_rn(mf[6], xf, den0), wb = __fmaf_rn(mf[6], xf + 3.f, den0);
    const float wmin = fminf(wa, wb);
    unsigned amb = 0xFu;
    if (wmin > 1e-3f) {
        const bool newton = fabsf(mf[6]) * 3.f < 1e-4f * wmin;   // uniform per frame in practice
        float r[4], fx[4], fy[4];
        r[0] = __fdividef(1.f, wa);
#pragma unroll
        for (int j = 1; j < 4; j++) {
            const float den = __fmaf_rn(mf[6], xf + (float)j, den0);
            r[j] = newton ? __fmaf_rn(r[j - 1], __fmaf_rn(-den, r[j - 1], 1.f), r[j - 1]) : __fdividef(1.f, den);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float xj = xf + (float)j;
            fx[j] = __fmaf_rn(mf[0], xj, nx0) * r[j];
            fy[j] = __fmaf_rn(mf[3], xj, ny0) * r[j];
        }
        const float lim_x = (float)p.sw + 0.25f, lim_y = (float)p.sh + 0.25f;
        const bool off = (fx[0] < -1.25f && fx[3] < -1.25f) || (fx[0] > lim_x && fx[3] > lim_x) ||
                         (fy[0] < -1.25f && fy[3] < -1.25f) || (fy[0] > lim_y && fy[3] > lim_y);
        amb = 0u;
        if (!off) {
            const float rmax = fmaxf(r[0], r[3]);
            const float magx = __fmaf_rn(fabsf(mf[0]), xf + 3.f, __fmaf_rn(fabsf(mf[1]), yf, fabsf(mf[2]))) * rmax;
            const float magy = __fmaf_rn(fabsf(mf[3]), xf + 3.f, __fmaf_rn(fabsf(mf[4]), yf, fabsf(mf[5]))) * rmax;
            const float thr_x = __fmaf_rn(48.f * 5.97e-8f, magx, 1e-6f), thr_y = __fmaf_rn(48.f * 5.97e-8f, magy, 1e-6f);
            const float kMagic = 12582912.f;   // 1.5 * 2^23
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float tx = fx[j] + kMagic, ty = fy[j] + kMagic;
                const float rx = tx - kMagic, ry = ty - kMagic;
                const int ix = __float_as_int(tx) - 0x4B400000, iy = __float_as_int(ty) - 0x4B400000;
                const bool sane = fabsf(fx[j]) < 2097152.f && fabsf(fy[j]) < 2097152.f;   // the trick is exact below 2^22
                const bool near_half = (0.5f - fabsf(fx[j] - rx) < thr_x) || (0.5f - fabsf(fy[j] - ry) < thr_y);
                if (sane && (unsigned)ix < (unsigned)p.sw && (unsigned)iy < (unsigned)p.sh) w[j] = __ldg(p.wimg + (iy * p.sw + ix));
                if (sane && near_half) amb |= 1u << j;   // (not sane = millions of px away from the frame: weight 0 for sure)
            }
        }
    }
    if (amb) {   // exact OpenCV arithmetic for the px the FP32 pass could not decide
        double M[9];
#pragma unroll
        for (int i = 0; i < 9; i++) M[i] = J.hinv[i];
        RowBase rb = row_base(M, x, y);
        const double x1 = (double)(x & 63);
        while (amb) {
            const int j = __ffs(amb) - 1;
            amb &= amb - 1;
            double fx, fy;
            px_coord(M, rb, x1 + (double)j, fx, fy);
            int nx = rnd(fx), ny = rnd(fy);
            float wv = ((unsigned)nx < (unsigned)p.sw && (unsigned)ny < (unsigned)p.sh) ? __ldg(p.wimg + (ny * p.sw + nx)) : 0.f;
            if (j == 0) w[0] = wv; else if (j == 1) w[1] = wv; else if (j == 2) w[2] = wv; else w[3] = wv;
        }
    }
    return make_float4(w[0], w[1], w[2], w[3]);
}
template <bool LEAN>
__device__ __forceinline__ float4 mbw_weights4_sel(const GroupParams& p, const FrameJob& J, int x, int y) {
    if constexpr (LEAN) return mbw_weights4_lean(p, J, x, y);
    else return mbw_weights4(p, J, x, y);
}

__global__ void __launch_bounds__(256) mbw_warp_kernel(const __grid_constant__ GroupParams p) {
    const FrameJob& J = p.jobs[blockIdx.y];
    const int ww = J.wnx * kEle, wh = J.wny * kEle;
    const int bpr = J.wnx;
    int by = blockIdx.x / bpr, bx = blockIdx.x - by * bpr;
    if (by * 4 >= wh) return;
    int u = bx * kEle + (threadIdx.x & 63) * 4, v = by * 4 + (threadIdx.x >> 6);
    int x = u + J.wx * kEle, y = v + J.wy * kEle;
    const float4 w = mbw_weights4(p, J, x, y);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.scratch + J.w_off[0]) + (size_t)v * ww + u) = w;
}
__global__ void __launch_bounds__(256) mbw_warp_lean_kernel(const __grid_constant__ GroupParams p) {   // EXPERIMENTAL, see mbw_weights4_lean
    const FrameJob& J = p.jobs[blockIdx.y];
    const int ww = J.wnx * kEle, wh = J.wny * kEle;
    const int bpr = J.wnx;
    int by = blockIdx.x / bpr, bx = blockIdx.x - by * bpr;
    if (by * 4 >= wh) return;
    int u = bx * kEle + (threadIdx.x & 63) * 4, v = by * 4 + (threadIdx.x >> 6);
    const float4 w = mbw_weights4_lean(p, J, u + J.wx * kEle, v + J.wy * kEle);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.scratch + J.w_off[0]) + (size_t)v * ww + u) = w;
}
cudaError_t launch_mbw_warp(const GroupParams& p, cudaStream_t stream, bool lean) {
    dim3 g(p.max_wnx * p.max_wny * (kEle / 4), p.n_frames);
    if (lean) mbw_warp_lean_kernel<<<g, 256, 0, stream>>>(p);
    else mbw_warp_kernel<<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---- 1a'. EXPERIMENTAL (M2D_WFUSED=1, not the default, not yet measured): weight warp + first weight pyrDown in one
// kernel.  A CTA computes a 128 x 32 block of level-0 weights plus the halo the 5-tap filter needs (136 x 35) into
// shared memory with mbw_weights4 (the FP32 pass makes the 16 % halo cheap), writes the block to the level-0 plane
// (mbs_decide reads it) and produces its 64 x 16 level-1 weights straight from shared memory, so level 0 is never read
// back.  Same tiling and border handling as mb_warp_pyr_kernel: the pyrDown taps are reflected (BORDER_REFLECT_101) in
// REGION coordinates, which always lands inside the block's own tile, so px outside the region are never needed.
template <bool LEAN>
__global__ void __launch_bounds__(256) mbw_warp_pyr_kernel(const __grid_constant__ GroupParams p) {
    __shared__ float sW[kFSH][kFSW];
    const FrameJob& J = p.jobs[blockIdx.y];
    const int ww = J.wnx * kEle, wh = J.wny * kEle, rw = J.nx * kEle, rh = J.ny * kEle;
    const int bpr = ww / kFW;
    int by = blockIdx.x / bpr, bx = blockIdx.x - by * bpr;
    if (by * kFH >= wh) return;
    const int x0 = bx * kFW + J.wx * kEle, y0 = by * kFH + J.wy * kEle;  // block origin, region coordinates
    float* W0 = reinterpret_cast<float*>(p.scratch + J.w_off[0]);
    constexpr int kGroups = (kFSW / 4) * kFSH;
    for (int gi = threadIdx.x; gi < kGroups; gi += 256) {
        int r = gi / (kFSW / 4), c = (gi - r * (kFSW / 4)) * 4;
        int x = x0 - 4 + c, y = y0 - 2 + r;
        if (x < 0 || x >= rw || y < 0 || y >= rh) continue;  // outside the region: reflected taps never read it
        const float4 w = mbw_weights4_sel<LEAN>(p, J, x, y);
        *reinterpret_cast<float4*>(&sW[r][c]) = w;
        if (c >= 4 && c < 4 + kFW && r >= 2 && r < 2 + kFH)   // the block itself goes to the level-0 plane
            *reinterpret_cast<float4*>(W0 + (size_t)(y - J.wy * kEle) * ww + (x - J.wx * kEle)) = w;
    }
    __syncthreads();
    const int dww = ww >> 1;
    float* W1 = reinterpret_cast<float*>(p.scratch + J.w_off[1]);
    const int ul = threadIdx.x & 63, vl0 = (threadIdx.x >> 6) * 4;
    const int U = (x0 >> 1) + ul, V0 = (y0 >> 1) + vl0;  // region coordinates at level 1
    int cs[5];
#pragma unroll
    for (int d = 0; d < 5; d++) cs[d] = reflect101_idx(2 * U + d - 2, rw) - (x0 - 4);
    float hw[5];
#pragma unroll
    for (int r = 0; r < 11; r++) {
        int rr = reflect101_idx(2 * V0 + r - 2, rh) - (y0 - 2);
        const float* wr = sW[rr];
        // f32, OpenCV 2.4.9 association: s0*6 + (s-1 + s1)*4 + s-2 + s2, left to right
        hw[r % 5] = wr[cs[2]] * 6.f + (wr[cs[1]] + wr[cs[3]]) * 4.f + wr[cs[0]] + wr[cs[4]];
        if (r >= 4 && (r & 1) == 0) {
            const int k = (r - 4) >> 1, i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
            float t0 = (hw[i0] + hw[i4]) + (hw[i2] + hw[i2]);
            float t1 = (hw[i1] + hw[i3]) + hw[i2];
            W1[(size_t)(V0 + k - ((J.wy * kEle) >> 1)) * dww + (U - ((J.wx * kEle) >> 1))] = (t0 + t1 * 4.f) * (1.f / 256.f);
        }
    }
}
cudaError_t launch_mbw_warp_pyr(const GroupParams& p, cudaStream_t stream, bool lean) {
    dim3 g(p.max_wnx * p.max_wny * (kEle / kFW) * (kEle / kFH), p.n_frames);
    if (lean) mbw_warp_pyr_kernel<true><<<g, 256, 0, stream>>>(p);
    else mbw_warp_pyr_kernel<false><<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---- 1b. weight pyrDown l -> l+1 (f32, OpenCV 2.4.9 association), same tiling as mb_pyrdown_kernel ----
__global__ void __launch_bounds__(256) mbw_pyrdown_kernel(const __grid_constant__ GroupParams p, int l) {
    const FrameJob& J = p.jobs[blockIdx.y];
    const int ns = kEle >> l, nd = kEle >> (l + 1);
    const int sww = J.wnx * ns, swh = J.wny * ns, srw = J.nx * ns, srh = J.ny * ns, sox = J.wx * ns, soy = J.wy * ns;
    const int dww = J.wnx * nd, dwh = J.wny * nd, dox = J.wx * nd, doy = J.wy * nd;
    const int bpr = (dww + 63) / 64;
    int by = blockIdx.x / bpr, bx = blockIdx.x - by * bpr;
    int u = (bx * 32 + threadIdx.x) * 2, v0 = (by * 8 + threadIdx.y) * 4;
    if (u >= dww || v0 >= dwh) return;
    const float* SW = reinterpret_cast<const float*>(p.scratch + J.w_off[l]);
    float* DW = reinterpret_cast<float*>(p.scratch + J.w_off[l + 1]);
    const bool two = (u + 1) < dww;
    int U = u + dox;
    int xs[7];
    int c0 = 2 * U - 2 - sox;
    const bool fast = (2 * U - 2 >= 0) && (2 * U + 4 < srw) && (c0 >= 0) && (c0 + 6 < sww);
    if (fast) {
#pragma unroll
        for (int d = 0; d < 7; d++) xs[d] = c0 + d;
    } else {
#pragma unroll
        for (int d = 0; d < 7; d++) xs[d] = clampi(reflect101_idx(2 * U + d - 2, srw) - sox, 0, sww - 1);
    }
    int V0 = v0 + doy;
    float h0[5], h1[5];
#pragma unroll
    for (int r = 0; r < 11; r++) {
        int ys = clampi(reflect101_idx(2 * V0 + r - 2, srh) - soy, 0, swh - 1);
        const float* wr = SW + (size_t)ys * sww;
        float f[7];
        if (fast) {
            const float2* w2 = reinterpret_cast<const float2*>(wr + xs[0]);
            float2 fa = w2[0], fb = w2[1], fc = w2[2];
            f[0] = fa.x; f[1] = fa.y; f[2] = fb.x; f[3] = fb.y; f[4] = fc.x; f[5] = fc.y; f[6] = wr[xs[0] + 6];
        } else {
#pragma unroll
            for (int d = 0; d < 7; d++) f[d] = wr[xs[d]];
        }
        h0[r % 5] = f[2] * 6.f + (f[1] + f[3]) * 4.f + f[0] + f[4];
        h1[r % 5] = f[4] * 6.f + (f[3] + f[5]) * 4.f + f[2] + f[6];
        if (r >= 4 && (r & 1) == 0) {
            int k = (r - 4) >> 1, v = v0 + k;
            if (v < dwh) {
                const int i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
                float t00 = (h0[i0] + h0[i4]) + (h0[i2] + h0[i2]), t10 = (h0[i1] + h0[i3]) + h0[i2];
                float t01 = (h1[i0] + h1[i4]) + (h1[i2] + h1[i2]), t11 = (h1[i1] + h1[i3]) + h1[i2];
                float ow0 = (t00 + t10 * 4.f) * (1.f / 256.f), ow1 = (t01 + t11 * 4.f) * (1.f / 256.f);
                size_t o = (size_t)v * dww + u;
                if (two && !(dww & 1)) *reinterpret_cast<float2*>(DW + o) = make_float2(ow0, ow1);
                else { DW[o] = ow0; if (two) DW[o + 1] = ow1; }
            }
        }
    }
}
cudaError_t launch_mbw_pyrdown(const GroupParams& p, int level, cudaStream_t stream) {
    int nd = kEle >> (level + 1);
    int blocks = ((p.max_wnx * nd + 63) / 64) * ((p.max_wny * nd + 31) / 32);
    dim3 b(32, 8), g(blocks, p.n_frames);
    mbw_pyrdown_kernel<<<g, b, 0, stream>>>(p, level);
    return cudaGetLastError();
}

// ---- 1c. weight pyramid tail (small deep levels, one CTA per frame) ----
__global__ void __launch_bounds__(1024) mbw_pyrtail_kernel(const __grid_constant__ GroupParams p, int l_first) {
    const FrameJob& J = p.jobs[blockIdx.x];
    for (int l = l_first; l + 1 < p.levels; l++) {
        const int ns = kEle >> l, nd = kEle >> (l + 1);
        const int sww = J.wnx * ns, swh = J.wny * ns, srw = J.nx * ns, srh = J.ny * ns, sox = J.wx * ns, soy = J.wy * ns;
        const int dww = J.wnx * nd, dwh = J.wny * nd, dox = J.wx * nd, doy = J.wy * nd;
        const float* SW = reinterpret_cast<const float*>(p.scratch + J.w_off[l]);
        float* DW = reinterpret_cast<float*>(p.scratch + J.w_off[l + 1]);
        for (int o = threadIdx.x; o < dww * dwh; o += blockDim.x) {
            int v = o / dww, u = o - v * dww;
            int U = u + dox, V = v + doy;
            int xs[5], ys[5];
#pragma unroll
            for (int d = 0; d < 5; d++) {
                xs[d] = clampi(reflect101_idx(2 * U + d - 2, srw) - sox, 0, sww - 1);
                ys[d] = clampi(reflect101_idx(2 * V + d - 2, srh) - soy, 0, swh - 1);
            }
            float hw[5];
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const float* wr = SW + (size_t)ys[r] * sww;
                hw[r] = wr[xs[2]] * 6.f + (wr[xs[1]] + wr[xs[3]]) * 4.f + wr[xs[0]] + wr[xs[4]];
            }
            float t0 = (hw[0] + hw[4]) + (hw[2] + hw[2]);
            float t1 = (hw[1] + hw[3]) + hw[2];
            DW[o] = (t0 + t1 * 4.f) * (1.f / 256.f);
        }
        __syncthreads();
    }
}
cudaError_t launch_mbw_pyrtail(const GroupParams& p, int l_first, cudaStream_t stream) {
    mbw_pyrtail_kernel<<<p.n_frames, 1024, 0, stream>>>(p, l_first);
    return cudaGetLastError();
}

// ---- 2. decide: the scan of mb_select_kernel without the Laplacian; winners go to the winner map + cell flags ----
__global__ void __launch_bounds__(256, 6) mbs_decide_kernel(const __grid_constant__ GroupParams p, const __grid_constant__ TileLayout lay) {
    const TileWork T = p.tiles[blockIdx.x];
    int q = blockIdx.y * 256 + threadIdx.x;
    int l = 0, n = kEle, half = kEle / 2;
    for (; l < p.levels; l++) {
        n = kEle >> l;
        half = n > 1 ? n / 2 : 1;
        int cnt = half * half;
        if (q < cnt) break;
        q -= cnt;
    }
    const bool valid = l < p.levels;
    if (!valid) { l = p.levels - 1; n = kEle >> l; half = n > 1 ? n / 2 : 1; q = 0; }
    const int qy = q / half, qx = q - qy * half;
    const int py = qy * 2, px = qx * 2;
    const bool quad = n > 1;
    const size_t to = (size_t)py * n + px;
    float* tw = reinterpret_cast<float*>(T.state + lay.wgt_off[l]) + to;
    float bw[4];
    if (T.fresh) { bw[0] = bw[1] = bw[2] = bw[3] = -INFINITY; }
    else if (quad) {
        float2 t0 = *reinterpret_cast<const float2*>(tw), t1 = *reinterpret_cast<const float2*>(tw + n);
        bw[0] = t0.x; bw[1] = t0.y; bw[2] = t1.x; bw[3] = t1.y;
    } else { bw[0] = tw[0]; bw[1] = bw[2] = bw[3] = 0.f; }
    int best[4] = {-1, -1, -1, -1};
    unsigned wins = 0;
    const int lane = threadIdx.x & 31;
    const bool uniform = __all_sync(0xffffffffu, l == __shfl_sync(0xffffffffu, l, 0));
    for (int c0 = 0; c0 < T.count; c0 += 32) {
        unsigned long long wb = 0ull;
        int stride = 0;
        if (uniform && c0 + lane < T.count) {
            const TileEntry E = p.entries[T.first + c0 + lane];
            const FrameJob& J = p.jobs[E.frame];
            stride = J.wnx * n;
            wb = reinterpret_cast<unsigned long long>(p.scratch + J.w_off[l]) +
                 4ull * ((size_t)((E.rty - J.wy) * n) * stride + (size_t)((E.rtx - J.wx) * n));
        }
        const int m = min(32, T.count - c0);
#pragma unroll 4
        for (int i = 0; i < m; i++) {
            const float* W;
            int st;
            if (uniform) {
                W = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, wb, i));
                st = __shfl_sync(0xffffffffu, stride, i);
            } else {
                const TileEntry E = p.entries[T.first + c0 + i];
                const FrameJob& J = p.jobs[E.frame];
                st = J.wnx * n;
                W = reinterpret_cast<const float*>(p.scratch + J.w_off[l]) + (size_t)((E.rty - J.wy) * n) * st + (size_t)((E.rtx - J.wx) * n);
            }
            const float* qp = W + (size_t)py * st + px;
            float s[4];
            if (quad) {
                float2 t0 = *reinterpret_cast<const float2*>(qp), t1 = *reinterpret_cast<const float2*>(qp + st);
                s[0] = t0.x; s[1] = t0.y; s[2] = t1.x; s[3] = t1.y;
            } else { s[0] = qp[0]; s[1] = s[2] = s[3] = -INFINITY; }
            const unsigned cw = !(T.fresh && (c0 + i) == 0);
#pragma unroll
            for (int k = 0; k < 4; k++)
                if ((quad || k == 0) && s[k] >= bw[k]) { bw[k] = s[k]; best[k] = c0 + i; wins += cw; }   // '>=' : MultiBandMap2DCPU.cpp:542
        }
    }
    if (!valid) { best[0] = best[1] = best[2] = best[3] = -1; wins = 0; }
    if (p.stats) {
        unsigned long long w = warp_sum(wins);
        if (lane == 0 && w) atomicAdd(p.stats + l, w);
    }
    if (!valid) return;
    // winner map: entry index inside the tile's list, 0xFFFF = the state stands
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
    // cell flags of the winning frames (a 2x2 quad sits inside one cell while cells are >= 2 px, i.e. l <= 4)
    size_t marked[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        marked[k] = ~(size_t)0;
        if (best[k] < 0) continue;
        const TileEntry E = p.entries[T.first + best[k]];
        const FrameJob& J = p.jobs[E.frame];
        int wpx = (E.rtx - J.wx) * n + px + (k & 1), wpy = (E.rty - J.wy) * n + py + (k >> 1);
        size_t idx = cell_base(p, E.frame, l) + (size_t)((wpy << l) >> 5) * (J.wnx * 8) + ((wpx << l) >> 5);
        bool dup = false;
#pragma unroll
        for (int k2 = 0; k2 < k; k2++) dup |= marked[k2] == idx;
        marked[k] = idx;
        if (!dup && !p.win[idx]) p.win[idx] = 1;
    }
}
cudaError_t launch_mbs_decide(const GroupParams& p, const TileLayout& lay, cudaStream_t stream) {
    if (p.n_tiles == 0) return cudaSuccess;
    int quads = 0;
    for (int l = 0; l < p.levels; l++) {
        int n = kEle >> l, half = n > 1 ? n / 2 : 1;
        quads += half * half;
    }
    dim3 g(p.n_tiles, (quads + 255) / 256);
    mbs_decide_kernel<<<g, 256, 0, stream>>>(p, lay);
    return cudaGetLastError();
}

// ---- 2'. EXPERIMENTAL (M2D_DCULL=1, not the default, never run on a GPU yet): decide with best-first order and
// bound-based culling.  "Largest weight, latest frame on ties" (what the sequential `>=` scan leaves behind, the
// pre-existing state counting as frame -1) is order-free, so the host sorts a tile's frames by footprint-centre distance
// and the kernel skips, without loading anything, every frame whose weight UPPER BOUND over the warp's px is below the
// warp's current minimum.  The bound (validated against real weight pyramids in tests/test_weights_first_host.py,
// level_weight_upper_bound): the level-0 support rect of the warp's px, mapped through the inverse homography by its 4
// corners, bounding box grown by 1 px; the radial weight at the box's point nearest to the frame centre bounds every
// sample, and pyrDown (a convex combination whose borders reflect inwards) cannot exceed it beyond float rounding.
// Lane i bounds frame c0+i, as in weighted_group_kernel.  Not used with collect_stats (the win counters follow the
// sequential semantics).
__device__ __forceinline__ float level_weight_upper_bound(const GroupParams& p, const float* __restrict__ m, float x0, float y0, float x1, float y1) {
    float bx0 = 3.0e38f, bx1 = -3.0e38f, by0 = 3.0e38f, by1 = -3.0e38f;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float x = (c & 1) ? x1 : x0, y = (c & 2) ? y1 : y0;
        const float den = m[6] * x + m[7] * y + m[8];
        if (!(den > 1e-3f)) return INFINITY;   // cannot reason about this frame here: never cull it
        const float r = 1.f / den;
        const float sx = (m[0] * x + m[1] * y + m[2]) * r, sy = (m[3] * x + m[4] * y + m[5]) * r;
        bx0 = fminf(bx0, sx); bx1 = fmaxf(bx1, sx); by0 = fminf(by0, sy); by1 = fmaxf(by1, sy);
    }
    bx0 -= 1.f; bx1 += 1.f; by0 -= 1.f; by1 += 1.f;
    if (bx1 < -0.5f || bx0 > (float)p.sw - 0.5f || by1 < -0.5f || by0 > (float)p.sh - 0.5f) return 0.f;   // samples outside the frame only
    const float xc = (float)(p.sw / 2), yc = (float)(p.sh / 2);
    const float dx = fmaxf(0.f, fmaxf(bx0 - xc, xc - bx1)), dy = fmaxf(0.f, fmaxf(by0 - yc, yc - by1));
    const float dis = 1.f - fminf(sqrtf(dx * dx + dy * dy) / sqrtf(xc * xc + yc * yc), 1.f);
    const float v = p.weight_type == 0 ? dis : dis * dis;
    return fmaxf(v, 1e-5f) * (1.f + 3e-5f) + 1e-7f;
}

__global__ void __launch_bounds__(256, 6) mbs_decide_bf_kernel(const __grid_constant__ GroupParams p, const __grid_constant__ TileLayout lay) {
    const TileWork T = p.tiles[blockIdx.x];
    int q = blockIdx.y * 256 + threadIdx.x;
    int l = 0, n = kEle, half = kEle / 2;
    for (; l < p.levels; l++) {
        n = kEle >> l;
        half = n > 1 ? n / 2 : 1;
        int cnt = half * half;
        if (q < cnt) break;
        q -= cnt;
    }
    const bool valid = l < p.levels;
    if (!valid) { l = p.levels - 1; n = kEle >> l; half = n > 1 ? n / 2 : 1; q = 0; }
    const int qy = q / half, qx = q - qy * half;
    const int py = qy * 2, px = qx * 2;
    const bool quad = n > 1;
    const size_t to = (size_t)py * n + px;
    float* tw = reinterpret_cast<float*>(T.state + lay.wgt_off[l]) + to;
    float bw[4];
    if (T.fresh) { bw[0] = bw[1] = bw[2] = bw[3] = -INFINITY; }
    else if (quad) {
        float2 t0 = *reinterpret_cast<const float2*>(tw), t1 = *reinterpret_cast<const float2*>(tw + n);
        bw[0] = t0.x; bw[1] = t0.y; bw[2] = t1.x; bw[3] = t1.y;
    } else { bw[0] = tw[0]; bw[1] = bw[2] = bw[3] = INFINITY; }   // px that do not exist never lower the warp minimum
    int best[4] = {-1, -1, -1, -1};    // position of the winner in the tile's (sorted) entry list
    int bestf[4] = {-1, -1, -1, -1};   // its frame index = feed order; -1 = the state, which loses every tie
    const int lane = threadIdx.x & 31;
    const bool uniform = __all_sync(0xffffffffu, valid && l == __shfl_sync(0xffffffffu, l, 0));
    // the warp's px set as a rect of level-l px of the tile, and its level-0 support
    const int pxlo = __reduce_min_sync(0xffffffffu, px), pxhi = __reduce_max_sync(0xffffffffu, px + (quad ? 1 : 0));
    const int pylo = __reduce_min_sync(0xffffffffu, py), pyhi = __reduce_max_sync(0xffffffffu, py + (quad ? 1 : 0));
    const int R = l == 0 ? 0 : (2 << l) - 2;
    for (int c0 = 0; c0 < T.count; c0 += 32) {
        unsigned long long wb = 0ull;
        int stride = 0, myframe = 0;
        float ub = INFINITY;
        if (c0 + lane < T.count) {
            const TileEntry E = p.entries[T.first + c0 + lane];
            const FrameJob& J = p.jobs[E.frame];
            myframe = E.frame;
            stride = J.wnx * n;
            wb = reinterpret_cast<unsigned long long>(p.scratch + J.w_off[l]) +
                 4ull * ((size_t)((E.rty - J.wy) * n) * stride + (size_t)((E.rtx - J.wx) * n));
            if (uniform) {
                const int ox = E.rtx * n, oy = E.rty * n;   // the tile's origin in the frame's region, level-l px
                ub = level_weight_upper_bound(p, J.hinvf, (float)(((ox + pxlo) << l) - R), (float)(((oy + pylo) << l) - R),
                                              (float)(((ox + pxhi) << l) + R), (float)(((oy + pyhi) << l) + R));
            }
        }
        float mine = fminf(fminf(bw[0], bw[1]), fminf(bw[2], bw[3]));
        float wmin = uniform ? __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(fmaxf(mine, 0.f)))) : -INFINITY;  // >= 0: int order == float order
        if (T.fresh && __any_sync(0xffffffffu, mine == -INFINITY)) wmin = -INFINITY;   // nothing decided yet somewhere in the warp
        unsigned mask = __ballot_sync(0xffffffffu, c0 + lane < T.count && ub >= wmin);
        while (mask) {
            const int i = __ffs(mask) - 1;
            mask &= mask - 1;
            const float* W = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, wb, i));
            const int st = __shfl_sync(0xffffffffu, stride, i);
            const int fr = __shfl_sync(0xffffffffu, myframe, i);
            const float* qp = W + (size_t)py * st + px;
            float s[4];
            if (quad) {
                float2 t0 = *reinterpret_cast<const float2*>(qp), t1 = *reinterpret_cast<const float2*>(qp + st);
                s[0] = t0.x; s[1] = t0.y; s[2] = t1.x; s[3] = t1.y;
            } else { s[0] = qp[0]; s[1] = s[2] = s[3] = -INFINITY; }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if ((quad || k == 0) && (s[k] > bw[k] || (s[k] == bw[k] && fr > bestf[k]))) { bw[k] = s[k]; best[k] = c0 + i; bestf[k] = fr; }
            if (uniform && mask) {   // the state only improves: re-filter what is still queued
                mine = fminf(fminf(bw[0], bw[1]), fminf(bw[2], bw[3]));
                const bool undecided = __any_sync(0xffffffffu, mine == -INFINITY);
                wmin = undecided ? -INFINITY : __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(fmaxf(mine, 0.f))));
                mask &= __ballot_sync(0xffffffffu, ub >= wmin);
            }
        }
    }
    if (!valid) return;
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
    size_t marked[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        marked[k] = ~(size_t)0;
        if (best[k] < 0) continue;
        const TileEntry E = p.entries[T.first + best[k]];
        const FrameJob& J = p.jobs[E.frame];
        int wpx = (E.rtx - J.wx) * n + px + (k & 1), wpy = (E.rty - J.wy) * n + py + (k >> 1);
        size_t idx = cell_base(p, E.frame, l) + (size_t)((wpy << l) >> 5) * (J.wnx * 8) + ((wpx << l) >> 5);
        bool dup = false;
#pragma unroll
        for (int k2 = 0; k2 < k; k2++) dup |= marked[k2] == idx;
        marked[k] = idx;
        if (!dup && !p.win[idx]) p.win[idx] = 1;
    }
}
cudaError_t launch_mbs_decide_bf(const GroupParams& p, const TileLayout& lay, cudaStream_t stream) {
    if (p.n_tiles == 0) return cudaSuccess;
    int quads = 0;
    for (int l = 0; l < p.levels; l++) {
        int n = kEle >> l, half = n > 1 ? n / 2 : 1;
        quads += half * half;
    }
    dim3 g(p.n_tiles, (quads + 255) / 256);
    mbs_decide_bf_kernel<<<g, 256, 0, stream>>>(p, lay);
    return cudaGetLastError();
}

// ---- 3. propagate: need[k] = cells within reach of a win.  A win of level m in cell c' needs G_k valid in cells
// [c' - reach_lo[m][k], c' + reach_hi[m][k]] (both axes): the host derives the table by interval arithmetic over the
// exact taps (make_reach_table).  One thread per (frame, level, cell).
__global__ void __launch_bounds__(256) mbs_propagate_kernel(const __grid_constant__ GroupParams p) {
    const int f = blockIdx.y;
    const FrameJob& J = p.jobs[f];
    const int cw = J.wnx * 8, ch = J.wny * 8, nc = cw * ch, L = p.levels;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= L * nc) return;
    const int k = i / nc, c = i - k * nc, cy = c / cw, cx = c - cy * cw;
    bool v = false;
    for (int m = max(k - 1, 0); m < L && !v; m++) {
        const int rl = p.reach_lo[m][k], rh = p.reach_hi[m][k];
        if (rl == 0xFF) continue;
        const uint8_t* w = p.win + cell_base(p, f, m);   // a frame's flags are a few KB: L1/L2 resident
        // c is required by a win in c' iff c' - rl <= c <= c' + rh  <=>  c - rh <= c' <= c + rl
        const int y0 = max(cy - rh, 0), y1 = min(cy + rl, ch - 1), x0 = max(cx - rh, 0), x1 = min(cx + rl, cw - 1);
        for (int y = y0; y <= y1 && !v; y++)
            for (int x = x0; x <= x1; x++) v |= w[y * cw + x] != 0;
    }
    p.need[cell_base(p, f, k) + c] = v ? 1 : 0;
    if (p.stats && v) {   // collect_stats: px of level k this frame has to compute (a cell is (32 >> k)^2 px, 1 px at level 5)
        const unsigned long long side = (unsigned long long)max(32 >> k, 1);
        atomicAdd(p.stats + 20 + k, side * side);
    }
}
cudaError_t launch_mbs_propagate(const GroupParams& p, cudaStream_t stream) {
    dim3 g((p.levels * p.cells_max + 255) / 256, p.n_frames);
    mbs_propagate_kernel<<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// Reach table of the weights-first variant.  A win of level m occupying cell c (level-m px [c*B, (c+1)*B - 1], B = 32 >> m)
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

// ---- 4a. image warp of the needed level-0 cells: one CTA = one 32 x 32 cell, one thread = 4 px ----
__global__ void __launch_bounds__(256) mbs_warp_kernel(const __grid_constant__ GroupParams p) {
    const FrameJob& J = p.jobs[blockIdx.y];
    const int cw = J.wnx * 8, ch = J.wny * 8;
    const int cy = blockIdx.x / cw, cx = blockIdx.x - cy * cw;
    if (cy >= ch) return;
    if (!p.need[cell_base(p, blockIdx.y, 0) + blockIdx.x]) return;
    const int ww = J.wnx * kEle;
    int u = cx * 32 + (threadIdx.x & 7) * 4, v = cy * 32 + (threadIdx.x >> 3);
    int x = u + J.wx * kEle, y = v + J.wy * kEle;
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; i++) M[i] = J.hinv[i];
    const RawSrc R = make_raw_src(J.raw, J.raw_stride, nullptr, p.sw, p.sh);
    uint32_t g[4];
    mb_sample4<false>(p, R, M, x, y, g, nullptr);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(p.scratch + J.g_off[0]) + (size_t)v * ww + u) = make_uint4(g[0], g[1], g[2], g[3]);
}
cudaError_t launch_mbs_warp(const GroupParams& p, cudaStream_t stream) {
    dim3 g(p.max_wnx * 8 * p.max_wny * 8, p.n_frames);
    mbs_warp_kernel<<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// true when any cell under the output patch [u, u+nu) x [v, v+nv) of level L is needed
__device__ __forceinline__ bool patch_needed(const uint8_t* need, int L, int u, int v, int nu, int nv, int cw, int ch) {
    int cxa = (u << L) >> 5, cxb = min(((u + nu - 1) << L) >> 5, cw - 1), cya = (v << L) >> 5, cyb = min(((v + nv - 1) << L) >> 5, ch - 1);
    bool any = false;
    for (int cy = cya; cy <= cyb; cy++)
        for (int cx = cxa; cx <= cxb; cx++) any |= need[cy * cw + cx] != 0;
    return any;
}

// ---- 4b. image pyrDown l -> l+1 in needed cells.  One thread = 2 x 4 outputs; a warp = 16 x 16 outputs (one cell of level 1) ----
__global__ void __launch_bounds__(256, 4) mbs_pyrdown_kernel(const __grid_constant__ GroupParams p, int l) {
    const FrameJob& J = p.jobs[blockIdx.y];
    const int ns = kEle >> l, nd = kEle >> (l + 1);
    const int sww = J.wnx * ns, swh = J.wny * ns, srw = J.nx * ns, srh = J.ny * ns, sox = J.wx * ns, soy = J.wy * ns;
    const int dww = J.wnx * nd, dwh = J.wny * nd, dox = J.wx * nd, doy = J.wy * nd;
    const int bpr = (dww + 15) / 16;
    int by = blockIdx.x / bpr, bx = blockIdx.x - by * bpr;
    int u = (bx * 8 + threadIdx.x) * 2, v0 = (by * 32 + threadIdx.y) * 4;
    if (u >= dww || v0 >= dwh) return;
    if (!patch_needed(p.need + cell_base(p, blockIdx.y, l + 1), l + 1, u, v0, 2, 4, J.wnx * 8, J.wny * 8)) return;
    const uint32_t* SG = reinterpret_cast<const uint32_t*>(p.scratch + J.g_off[l]);
    uint32_t* DG = reinterpret_cast<uint32_t*>(p.scratch + J.g_off[l + 1]);
    const bool two = (u + 1) < dww;
    int U = u + dox;
    int xs[7];
    int c0 = 2 * U - 2 - sox;
    const bool fast = (2 * U - 2 >= 0) && (2 * U + 4 < srw) && (c0 >= 0) && (c0 + 6 < sww);
    if (fast) {
#pragma unroll
        for (int d = 0; d < 7; d++) xs[d] = c0 + d;
    } else {
#pragma unroll
        for (int d = 0; d < 7; d++) xs[d] = clampi(reflect101_idx(2 * U + d - 2, srw) - sox, 0, sww - 1);
    }
    int V0 = v0 + doy;
    uint32_t hb0[5], hg0[5], hb1[5], hg1[5];
#pragma unroll
    for (int r = 0; r < 11; r++) {
        int ys = clampi(reflect101_idx(2 * V0 + r - 2, srh) - soy, 0, swh - 1);
        const uint32_t* gr = SG + (size_t)ys * sww;
        uint32_t e[7];
        if (fast) {
            const uint2* g2 = reinterpret_cast<const uint2*>(gr + xs[0]);
            uint2 a = g2[0], b = g2[1], c = g2[2];
            e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[4] = c.x; e[5] = c.y; e[6] = gr[xs[0] + 6];
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
            int k = (r - 4) >> 1, v = v0 + k;
            if (v < dwh) {
                const int i0 = (r - 4) % 5, i1 = (r - 3) % 5, i2 = (r - 2) % 5, i3 = (r - 1) % 5, i4 = r % 5;
                uint32_t vbr0 = hb0[i0] + hb0[i4] + (hb0[i1] + hb0[i3]) * 4u + hb0[i2] * 6u;
                uint32_t vbr1 = hb1[i0] + hb1[i4] + (hb1[i1] + hb1[i3]) * 4u + hb1[i2] * 6u;
                uint32_t vg0 = hg0[i0] + hg0[i4] + (hg0[i1] + hg0[i3]) * 4u + hg0[i2] * 6u;
                uint32_t vg1 = hg1[i0] + hg1[i4] + (hg1[i1] + hg1[i3]) * 4u + hg1[i2] * 6u;
                uint32_t o0 = (((vbr0 + 0x00800080u) >> 8) & kM2) | (((vg0 + 128u) >> 8) << 8);
                uint32_t o1 = (((vbr1 + 0x00800080u) >> 8) & kM2) | (((vg1 + 128u) >> 8) << 8);
                size_t o = (size_t)v * dww + u;
                if (two && !(dww & 1)) *reinterpret_cast<uint2*>(DG + o) = make_uint2(o0, o1);
                else { DG[o] = o0; if (two) DG[o + 1] = o1; }
            }
        }
    }
}
cudaError_t launch_mbs_pyrdown(const GroupParams& p, int level, cudaStream_t stream) {
    int nd = kEle >> (level + 1);
    int blocks = ((p.max_wnx * nd + 15) / 16) * ((p.max_wny * nd + 127) / 128);
    dim3 b(8, 32), g(blocks, p.n_frames);
    mbs_pyrdown_kernel<<<g, b, 0, stream>>>(p, level);
    return cudaGetLastError();
}

// ---- 4c. image pyramid tail in needed cells ----
__global__ void __launch_bounds__(1024) mbs_pyrtail_kernel(const __grid_constant__ GroupParams p, int l_first) {
    const FrameJob& J = p.jobs[blockIdx.x];
    const int cw = J.wnx * 8, ch = J.wny * 8;
    for (int l = l_first; l + 1 < p.levels; l++) {
        const int ns = kEle >> l, nd = kEle >> (l + 1);
        const int sww = J.wnx * ns, swh = J.wny * ns, srw = J.nx * ns, srh = J.ny * ns, sox = J.wx * ns, soy = J.wy * ns;
        const int dww = J.wnx * nd, dwh = J.wny * nd, dox = J.wx * nd, doy = J.wy * nd;
        const uint32_t* SG = reinterpret_cast<const uint32_t*>(p.scratch + J.g_off[l]);
        uint32_t* DG = reinterpret_cast<uint32_t*>(p.scratch + J.g_off[l + 1]);
        const uint8_t* need = p.need + cell_base(p, blockIdx.x, l + 1);
        for (int o = threadIdx.x; o < dww * dwh; o += blockDim.x) {
            int v = o / dww, u = o - v * dww;
            if (!patch_needed(need, l + 1, u, v, 1, 1, cw, ch)) continue;
            int U = u + dox, V = v + doy;
            int xs[5], ys[5];
#pragma unroll
            for (int d = 0; d < 5; d++) {
                xs[d] = clampi(reflect101_idx(2 * U + d - 2, srw) - sox, 0, sww - 1);
                ys[d] = clampi(reflect101_idx(2 * V + d - 2, srh) - soy, 0, swh - 1);
            }
            uint32_t hbr[5], hg[5];
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t* gr = SG + (size_t)ys[r] * sww;
                uint32_t a = gr[xs[0]], b = gr[xs[1]], c = gr[xs[2]], d = gr[xs[3]], e = gr[xs[4]];
                hbr[r] = (c & kM2) * 6u + ((b & kM2) + (d & kM2)) * 4u + (a & kM2) + (e & kM2);
                hg[r] = ((c >> 8) & 0xFFu) * 6u + (((b >> 8) & 0xFFu) + ((d >> 8) & 0xFFu)) * 4u + ((a >> 8) & 0xFFu) + ((e >> 8) & 0xFFu);
            }
            uint32_t vbr = hbr[0] + hbr[4] + (hbr[1] + hbr[3]) * 4u + hbr[2] * 6u;
            uint32_t vg = hg[0] + hg[4] + (hg[1] + hg[3]) * 4u + hg[2] * 6u;
            DG[o] = (((vbr + 0x00800080u) >> 8) & kM2) | (((vg + 128u) >> 8) << 8);
        }
        __syncthreads();
    }
}
cudaError_t launch_mbs_pyrtail(const GroupParams& p, int l_first, cudaStream_t stream) {
    mbs_pyrtail_kernel<<<p.n_frames, 1024, 0, stream>>>(p, l_first);
    return cudaGetLastError();
}

// ---- 5. Laplacian of the winners (read from the winner map) into the tile state ----
__global__ void __launch_bounds__(256, 4) mbs_lap_kernel(const __grid_constant__ GroupParams p, const __grid_constant__ TileLayout lay) {
    const TileWork T = p.tiles[blockIdx.x];
    int q = blockIdx.y * 256 + threadIdx.x;
    int l = 0, n = kEle, half = kEle / 2;
    for (; l < p.levels; l++) {
        n = kEle >> l;
        half = n > 1 ? n / 2 : 1;
        int cnt = half * half;
        if (q < cnt) break;
        q -= cnt;
    }
    if (l >= p.levels) return;
    const int qy = q / half, qx = q - qy * half;
    const int py = qy * 2, px = qx * 2;
    const bool quad = n > 1;
    const size_t to = (size_t)py * n + px;
    const uint16_t* wm = p.wmap + (size_t)blockIdx.x * p.wmap_stride + lay.px_off[l] + to;
    int best[4] = {-1, -1, -1, -1};
    if (quad) {
        ushort2 a = *reinterpret_cast<const ushort2*>(wm), b = *reinterpret_cast<const ushort2*>(wm + n);
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
    int quads = 0;
    for (int l = 0; l < p.levels; l++) {
        int n = kEle >> l, half = n > 1 ? n / 2 : 1;
        quads += half * half;
    }
    dim3 g(p.n_tiles, (quads + 255) / 256);
    mbs_lap_kernel<<<g, 256, 0, stream>>>(p, lay);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// final tile gather of a sharded run: pack the raw state of n tiles into one contiguous buffer (or back)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_copy_kernel(uint8_t* const* __restrict__ tiles, uint8_t* __restrict__ buf, size_t tile_bytes, int to_buf) {
    const size_t n16 = tile_bytes / 16;
    uint4* t = reinterpret_cast<uint4*>(tiles[blockIdx.x]);
    uint4* b = reinterpret_cast<uint4*>(buf + (size_t)blockIdx.x * tile_bytes);
    for (size_t i = (size_t)blockIdx.y * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.y * 256) {
        if (to_buf) b[i] = t[i]; else t[i] = b[i];
    }
}
cudaError_t launch_tile_copy(uint8_t* const* d_tiles, int n, uint8_t* buf, size_t tile_bytes, int to_buf, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    dim3 g(n, 32);
    tile_copy_kernel<<<g, 256, 0, stream>>>(d_tiles, buf, tile_bytes, to_buf);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// collapse (MultiBandMap2DCPU::save, :779-841): paste tiles into per-level mosaics, restore from the Laplacian
// pyramid coarse -> fine (pyrUp + saturating add), convert to 8-bit and paint the background where weight == 0.
// ---------------------------------------------------------------------------------------------------------
// Weighted save(): paste every touched 256x256 BGRA tile into the zero-initialised mosaic (16 bytes per thread).
__global__ void __launch_bounds__(256) bgra_paste_kernel(const PasteItem* __restrict__ items, uint32_t* __restrict__ mosaic, int mosaic_w) {
    const PasteItem it = items[blockIdx.x];
    int i = blockIdx.y * 256 + threadIdx.x;          // 16-byte chunk index inside the tile: 64 per row, 256 rows
    int y = i >> 6, x = (i & 63) * 4;
    uint4 v = reinterpret_cast<const uint4*>(it.tile)[i];
    *reinterpret_cast<uint4*>(mosaic + (size_t)(it.ty * kEle + y) * mosaic_w + (size_t)it.tx * kEle + x) = v;
}
cudaError_t launch_bgra_paste(const PasteItem* d_items, int n_items, uint32_t* mosaic, int mosaic_w, cudaStream_t stream) {
    if (n_items == 0) return cudaSuccess;
    dim3 g(n_items, kEle * kEle / 4 / 256);
    bgra_paste_kernel<<<g, 256, 0, stream>>>(d_items, mosaic, mosaic_w);
    return cudaGetLastError();
}

// One launch pastes every touched tile (all levels) into the zero-initialised per-level mosaics.
__global__ void __launch_bounds__(256) mosaic_paste_kernel(const PasteItem* __restrict__ items, const __grid_constant__ TileLayout lay,
                                                           const __grid_constant__ MosaicSet ms) {
    const PasteItem it = items[blockIdx.x];
    int i = blockIdx.y * 256 + threadIdx.x;
    if (i >= lay.px_off[lay.levels]) return;
    int l = 0;
    while (i >= lay.px_off[l + 1]) l++;
    i -= lay.px_off[l];
    const int n = kEle >> l;
    int y = i / n, x = i - y * n;
    const int16_t* tl = reinterpret_cast<const int16_t*>(it.tile + lay.lap_off[l]);
    const MosaicLevel& m = ms.lv[l];
    size_t o = (size_t)(it.ty * n + y) * m.w + (size_t)it.tx * n + x;
    size_t plane = (size_t)n * n;
    m.g[0][o] = tl[i]; m.g[1][o] = tl[plane + i]; m.g[2][o] = tl[2 * plane + i];
    if (l == 0) ms.w0[o] = reinterpret_cast<const float*>(it.tile + lay.wgt_off[0])[i];
}
cudaError_t launch_mosaic_paste(const PasteItem* d_items, int n_items, const TileLayout& lay, const MosaicSet& ms, cudaStream_t stream) {
    if (n_items == 0) return cudaSuccess;
    dim3 g(n_items, (lay.px_off[lay.levels] + 255) / 256);
    mosaic_paste_kernel<<<g, 256, 0, stream>>>(d_items, lay, ms);
    return cudaGetLastError();
}
// Display-time collapse of ONE tile (MultiBandMap2DCPUEle::blend, MultiBandMap2DCPU.cpp:77-146): paste sub-rectangles of
// the tile and its 8 neighbours into a bordered per-level pyramid, then the usual restore chain, then crop + mask.
__global__ void __launch_bounds__(256) sub_paste_kernel(const SubPaste* __restrict__ items, const __grid_constant__ TileLayout lay,
                                                        const __grid_constant__ MosaicSet ms) {
    const SubPaste it = items[blockIdx.x];
    int i = blockIdx.y * 256 + threadIdx.x;
    if (i >= it.w * it.h) return;
    int y = i / it.w, x = i - y * it.w;
    const int n = kEle >> it.level;
    const int16_t* tl = reinterpret_cast<const int16_t*>(it.tile + lay.lap_off[it.level]);
    size_t s = (size_t)(it.sy + y) * n + (it.sx + x), plane = (size_t)n * n;
    const MosaicLevel& m = ms.lv[it.level];
    size_t o = (size_t)(it.dy + y) * m.w + (it.dx + x);
    m.g[0][o] = tl[s]; m.g[1][o] = tl[plane + s]; m.g[2][o] = tl[2 * plane + s];
}
cudaError_t launch_sub_paste(const SubPaste* d_items, int n_items, const TileLayout& lay, const MosaicSet& ms, cudaStream_t stream) {
    if (n_items == 0) return cudaSuccess;
    dim3 g(n_items, kEle * kEle / 256);
    sub_paste_kernel<<<g, 256, 0, stream>>>(d_items, lay, ms);
    return cudaGetLastError();
}
__global__ void tile_crop_kernel(MosaicLevel m0, int border, const float* __restrict__ w0, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kEle * kEle) return;
    int y = i / kEle, x = i - y * kEle;
    size_t o = (size_t)(y + border) * m0.w + (x + border);
    bool keep = w0[i] != 0.f;  // result.setTo(0, weights[0]==0), then convertTo(CV_8UC3)
    out[3 * i] = keep ? (uint8_t)min(max((int)m0.g[0][o], 0), 255) : 0;
    out[3 * i + 1] = keep ? (uint8_t)min(max((int)m0.g[1][o], 0), 255) : 0;
    out[3 * i + 2] = keep ? (uint8_t)min(max((int)m0.g[2][o], 0), 255) : 0;
}
cudaError_t launch_tile_crop(MosaicLevel m0, int border, const float* w0, uint8_t* out, cudaStream_t stream) {
    tile_crop_kernel<<<kEle * kEle / 256, 256, 0, stream>>>(m0, border, w0, out);
    return cudaGetLastError();
}

// fine += pyrUp(coarse), saturating int16 (restoreImageFromLaplacePyr)
__global__ void __launch_bounds__(256) mosaic_upadd_kernel(MosaicLevel C, MosaicLevel F) {
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= F.w || y >= F.h) return;
    int i = x >> 1, j = y >> 1;
    int c0 = pyrup_axis_lo(i - 1, C.w), c2 = pyrup_axis_hi(i + 1, C.w);
    int r0 = pyrup_axis_lo(j - 1, C.h), r2 = pyrup_axis_hi(j + 1, C.h);
    bool xodd = x & 1, yodd = y & 1;
    size_t o = (size_t)y * F.w + x;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int16_t* G = C.g[c];
        const int16_t *q0 = G + (size_t)r0 * C.w, *q1 = G + (size_t)j * C.w, *q2 = G + (size_t)r2 * C.w;
        int h0 = xodd ? (q0[i] + q0[c2]) * 4 : (q0[c0] + q0[i] * 6 + q0[c2]);
        int h1 = xodd ? (q1[i] + q1[c2]) * 4 : (q1[c0] + q1[i] * 6 + q1[c2]);
        int h2 = xodd ? (q2[i] + q2[c2]) * 4 : (q2[c0] + q2[i] * 6 + q2[c2]);
        int v = yodd ? (h1 + h2) * 4 : (h0 + h1 * 6 + h2);
        int up = sat16((v + 32) >> 6);
        F.g[c][o] = (int16_t)sat16((int)F.g[c][o] + up);
    }
}
cudaError_t launch_mosaic_upadd(MosaicLevel coarse, MosaicLevel fine, cudaStream_t stream) {
    dim3 b(32, 8), g((fine.w + 31) / 32, (fine.h + 7) / 8);
    mosaic_upadd_kernel<<<g, b, 0, stream>>>(coarse, fine);
    return cudaGetLastError();
}
__global__ void mosaic_final_kernel(MosaicLevel m, const float* __restrict__ w0, int background, uint8_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)m.w * m.h;
    if (i >= n) return;
    if (w0[i] == 0.f) {
        uint8_t bg = (uint8_t)min(max(background, 0), 255);
        out[3 * i] = bg; out[3 * i + 1] = bg; out[3 * i + 2] = bg;
    } else {
        out[3 * i] = (uint8_t)min(max((int)m.g[0][i], 0), 255);
        out[3 * i + 1] = (uint8_t)min(max((int)m.g[1][i], 0), 255);
        out[3 * i + 2] = (uint8_t)min(max((int)m.g[2][i], 0), 255);
    }
}
cudaError_t launch_mosaic_final(MosaicLevel m0, const float* w0, int background, uint8_t* out_bgr, cudaStream_t stream) {
    size_t n = (size_t)m0.w * m0.h;
    mosaic_final_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(m0, w0, background, out_bgr);
    return cudaGetLastError();
}

}  // namespace m2d
