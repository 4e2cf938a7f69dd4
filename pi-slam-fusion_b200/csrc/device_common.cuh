// device_common.cuh — device helpers shared by kernels.cu (weighted mode, dense multi-band pipeline, collapse) and
// kernels_wf.cu (weights-first multi-band pipeline).  Exact OpenCV arithmetic, see kernels.cu header.
#pragma once
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
// shared sampling helpers
// ---------------------------------------------------------------------------------------------------------
// Row base of cv::warpPerspectiveInvoker for the 64-px block containing x: X0 = M0*xb + M1*y + M2, etc.
// When M6 == 0 (an affine or row-affine homography: every exactly nadir frame) the denominator W0 + M6*x1 is W0 itself for
// every px of the row, bit for bit, so its reciprocal is taken ONCE per row segment instead of once per px.
struct RowBase { double X0, Y0, W0, Winv; bool const_w; };
__device__ __forceinline__ RowBase row_base(const double* M, int x, int y) {
    int xb = x & ~63;
    RowBase r;
    r.X0 = M[0] * xb + M[1] * y + M[2];
    r.Y0 = M[3] * xb + M[4] * y + M[5];
    r.W0 = M[6] * xb + M[7] * y + M[8];
    r.const_w = M[6] == 0.0;
    r.Winv = 0.0;
    if (r.const_w) r.Winv = (r.W0 != 0.0) ? 1.0 / r.W0 : 0.0;
    return r;
}
// Un-quantised source coordinate of the px at offset x1 (as a double, exactly the int->double value OpenCV
// multiplies by) inside the block.  INTER_LINEAR rounds 32*f, INTER_NEAREST rounds f (32/W == 32*(1/W) exactly).
__device__ __forceinline__ void px_coord(const double* M, const RowBase& r, double x1, double& fx, double& fy) {
    double W;
    if (r.const_w) W = r.Winv;
    else {
        W = r.W0 + M[6] * x1;
        W = (W != 0.0) ? 1.0 / W : 0.0;
    }
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

// f32 pyrDown association (oracle: pyr_down_f32).  mode 0 = OpenCV 2.4.9: rows s0*6 + (s-1 + s1)*4 + s-2 + s2 left to right,
// columns ((r0+r4)+(r2+r2)) + ((r1+r3)+r2)*4, scaled by 1/256 (PyrDownVec_32f: 8 columns per step, the remaining ocols % 8
// columns use the scalar r2*6 + (r1+r3)*4 + r0 + r4 -- only levels narrower than 8 px per tile and Map2DRender's sub-images
// have such a tail).  mode 1 = OpenCV 4.x, whose SIMD bodies
// cover output columns [1, hvec_end) horizontally -- s0*6 + ((s-1+s1)*4 + (s-2+s2)) -- and [0, vvec_end) vertically (same
// expression as 2.4.9), the remaining columns using the scalar expressions.  U = output column in REGION coordinates.
struct F32Assoc { int mode, hvec_end, vvec_end; };
__device__ __forceinline__ F32Assoc f32_assoc(int mode, int src_cols) {
    F32Assoc a;
    a.mode = mode;
    const int ocols = (src_cols + 1) / 2, width0 = min((src_cols - 3) / 2 + 1, ocols);
    a.hvec_end = (mode == 1 && width0 > 1) ? 1 + ((width0 - 1) / 4) * 4 : 0;
    a.vvec_end = (mode == 1) ? (ocols / 4) * 4 : (ocols / 8) * 8;   // 2.4.9's PyrDownVec_32f takes 8 columns per step, scalar tail
    return a;
}
__device__ __forceinline__ float pyr_h(const F32Assoc& a, int U, float sm2, float sm1, float s0, float sp1, float sp2) {
    if (a.mode == 1 && U >= 1 && U < a.hvec_end) return s0 * 6.f + ((sm1 + sp1) * 4.f + (sm2 + sp2));
    return s0 * 6.f + (sm1 + sp1) * 4.f + sm2 + sp2;
}
__device__ __forceinline__ float pyr_v(const F32Assoc& a, int U, float r0, float r1, float r2, float r3, float r4) {
    if (U < a.vvec_end) {
        const float t0 = (r0 + r4) + (r2 + r2), t1 = (r1 + r3) + r2;
        return (t0 + t1 * 4.f) * (1.f / 256.f);
    }
    return (r2 * 6.f + (r1 + r3) * 4.f + r0 + r4) * (1.f / 256.f);
}

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

}  // namespace m2d
