// kernels_render.cu — Map2DRender (Map2D type 4): the batch blender of Map2DRender.cpp:479-760 with its
// inline MultiBandBlender (:52-310), as sm_100a kernels.  Same exact-arithmetic rules as kernels.cu (--fmad=false).
//
// The reference warps every frame into ITS OWN bounding box (not the tile grid), pads it by reflection into a sub-image
// aligned to 1 << num_bands on the canvas, builds a Laplacian + weight pyramid of the sub-image and updates the canvas
// pyramids frame after frame.  Here:
//   rnd_warp      (frame, px)   level 0 of every frame's sub-image in one launch: the 8UC3 bilinear warp (fixed-point,
//                               BORDER_REFLECT) evaluated THROUGH copyMakeBorder's reflection -- the warped image itself is
//                               never stored -- and the 8-bit weight (nearest, constant 0) as f32 (/255) or s16 (+1)
//   rnd_pyrdown   (frame, px)   one level of every frame's Gaussian (packed u8x4 lanes) and weight pyramid
//   rnd_blend     (canvas px)   CANVAS-centric: every canvas px of every level walks the frames that cover it in feed
//                               order, so the canvas is read and written once per batch; the selection blend forms the
//                               Laplacian G_l - pyrUp(G_{l+1}) for the winning frame only, the weighted sums for every
//                               covering frame
//   rnd_normalize (canvas px)   normalizeUsingWeightMap (weighted sums only)
//   mosaic_upadd  (kernels.cu)  restoreImageFromLaplacePyr
//   rnd_final     (canvas px)   mask = weight > WEIGHT_EPS, masked px zeroed, crop, 16S -> 8U
#include <algorithm>
#include <type_traits>

#include "kernels.cuh"
#include "device_common.cuh"

namespace m2d {

// The weight image of Map2DRender::renderFrames (Map2DRender.cpp:507-529): centre at w*0.5, float arithmetic.
__global__ void rnd_weight_image_kernel(int sw, int sh, uint8_t* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= sw || i >= sh) return;
    const float x_center = (float)((double)sw * 0.5), y_center = (float)((double)sh * 0.5);
    const float dis_max_inv = (float)(1.0 / (double)sqrtf(x_center * x_center + y_center * y_center));
    float dis = ((float)i - y_center) * ((float)i - y_center) + ((float)j - x_center) * ((float)j - x_center);
    dis = 1.f - sqrtf(dis) * dis_max_inv;
    int p = (int)(dis * dis * 254.f) & 255;
    if (p < 1) p = 1;
    out[(size_t)i * sw + j] = (uint8_t)p;
}
cudaError_t launch_rnd_weight_image(int sw, int sh, uint8_t* out, cudaStream_t stream) {
    dim3 b(32, 8), g((sw + 31) / 32, (sh + 7) / 8);
    rnd_weight_image_kernel<<<g, b, 0, stream>>>(sw, sh, out);
    return cudaGetLastError();
}

// ---- level 0 of the sub-images ----
template <typename WT>
__global__ void __launch_bounds__(256) rnd_warp_kernel(const RenderJob* __restrict__ jobs, uint8_t* __restrict__ scratch,
                                                       const uint8_t* __restrict__ wimg, int src_w, int src_h) {
    const RenderJob& J = jobs[blockIdx.y];
    const int n = J.sw * J.sh;
    double M[9];
#pragma unroll
    for (int k = 0; k < 9; k++) M[k] = J.hinv[k];
    const RawSrc R = make_raw_src(J.raw, J.raw_stride, nullptr, src_w, src_h);
    uint32_t* G = reinterpret_cast<uint32_t*>(scratch + J.g_off[0]);
    WT* W = reinterpret_cast<WT*>(scratch + J.w_off[0]);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int v = i / J.sw, u = i - v * J.sw;
        const int ux = u - J.left, vy = v - J.top;
        const bool inside = (unsigned)ux < (unsigned)J.iw && (unsigned)vy < (unsigned)J.ih;
        const int wx = reflect_idx(ux, J.iw), wy = reflect_idx(vy, J.ih);   // copyMakeBorder(BORDER_REFLECT)
        const RowBase rb = row_base(M, wx, wy);
        double fx, fy;
        px_coord(M, rb, (double)(wx & 63), fx, fy);
        int X = rnd(fx * 32.0), Y = rnd(fy * 32.0);
        int nx = rnd(fx), ny = rnd(fy);
        if (__builtin_expect((unsigned)(X + 1048544) >= 2097088u || (unsigned)(Y + 1048544) >= 2097088u, 0)) {
            nx = sat_s16(nx); ny = sat_s16(ny);
            X = (sat_s16(X >> 5) << 5) | (X & 31); Y = (sat_s16(Y >> 5) << 5) | (Y & 31);
        }
        const int sx = X >> 5, sy = Y >> 5;
        uint32_t v00, v01, v10, v11;
        if ((unsigned)sx < (unsigned)(src_w - 1) && (unsigned)sy < (unsigned)(src_h - 1)) {
            raw_tap_pair_bgr(R, sx, sy, v00, v01);
            raw_tap_pair_bgr(R, sx, sy + 1, v10, v11);
        } else {
            int sx0 = reflect_once(sx, src_w), sx1 = reflect_once(sx + 1, src_w), sy0 = reflect_once(sy, src_h), sy1 = reflect_once(sy + 1, src_h);
            v00 = raw_tap_bgr(R, sx0, sy0); v01 = raw_tap_bgr(R, sx1, sy0); v10 = raw_tap_bgr(R, sx0, sy1); v11 = raw_tap_bgr(R, sx1, sy1);
        }
        // remapBilinear<FixedPtCast<int,uchar,15>>: (sum of v*w*32 + 16384) >> 15 == (sum of v*w + 512) >> 10, w = products of 5-bit fractions
        const uint32_t a = (uint32_t)(X & 31), b = (uint32_t)(Y & 31), wa0 = 32 - a, wb0 = 32 - b;
        const uint32_t br0 = (v00 & kM2) * wa0 + (v01 & kM2) * a, br1 = (v10 & kM2) * wa0 + (v11 & kM2) * a;
        const uint32_t g0 = ((v00 >> 8) & 0xFFu) * wa0 + ((v01 >> 8) & 0xFFu) * a, g1 = ((v10 >> 8) & 0xFFu) * wa0 + ((v11 >> 8) & 0xFFu) * a;
        const uint32_t B = ((br0 & 0xFFFFu) * wb0 + (br1 & 0xFFFFu) * b + 512u) >> 10;
        const uint32_t Rr = ((br0 >> 16) * wb0 + (br1 >> 16) * b + 512u) >> 10;
        const uint32_t Gg = (g0 * wb0 + g1 * b + 512u) >> 10;
        G[i] = B | (Gg << 8) | (Rr << 16);
        int m = 0;
        if (inside && (unsigned)nx < (unsigned)src_w && (unsigned)ny < (unsigned)src_h) m = __ldg(wimg + (size_t)ny * src_w + nx);
        if constexpr (sizeof(WT) == 4) W[i] = (WT)((float)m * (float)(1. / 255.));   // mask.convertTo(CV_32F, 1./255.)
        else W[i] = (WT)(m + (m != 0));                                              // convertTo(CV_16S); add 1 where mask != 0
    }
}
cudaError_t launch_rnd_warp(const RenderJob* d_jobs, int n_frames, int max_px, uint8_t* scratch, const uint8_t* wimg, int src_w, int src_h,
                            int s16_weights, cudaStream_t stream) {
    if (n_frames == 0) return cudaSuccess;
    dim3 g((unsigned)std::min((max_px + 255) / 256, 4096), n_frames);
    if (s16_weights) rnd_warp_kernel<int16_t><<<g, 256, 0, stream>>>(d_jobs, scratch, wimg, src_w, src_h);
    else rnd_warp_kernel<float><<<g, 256, 0, stream>>>(d_jobs, scratch, wimg, src_w, src_h);
    return cudaGetLastError();
}

// ---- pyrDown l -> l+1 of every frame's sub-image: BORDER_REFLECT_101 at the sub-image's own edges ----
template <typename WT>
__global__ void __launch_bounds__(256) rnd_pyrdown_kernel(const RenderJob* __restrict__ jobs, uint8_t* __restrict__ scratch, int l, int f32_mode) {
    const RenderJob& J = jobs[blockIdx.y];
    const int sws = J.sw >> l, shs = J.sh >> l, dw = sws >> 1, dh = shs >> 1;   // exact halves: sw, sh are multiples of 1 << num_bands
    const uint32_t* SG = reinterpret_cast<const uint32_t*>(scratch + J.g_off[l]);
    const WT* SW = reinterpret_cast<const WT*>(scratch + J.w_off[l]);
    uint32_t* DG = reinterpret_cast<uint32_t*>(scratch + J.g_off[l + 1]);
    WT* DW = reinterpret_cast<WT*>(scratch + J.w_off[l + 1]);
    const F32Assoc fa = f32_assoc(f32_mode, sws);
    for (int o = blockIdx.x * 256 + threadIdx.x; o < dw * dh; o += gridDim.x * 256) {
        const int v = o / dw, u = o - v * dw;
        int xs[5], ys[5];
#pragma unroll
        for (int d = 0; d < 5; d++) { xs[d] = reflect101_idx(2 * u + d - 2, sws); ys[d] = reflect101_idx(2 * v + d - 2, shs); }
        uint32_t hbr[5], hg[5];
        float hw[5];
        int hs[5];
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint32_t* gr = SG + (size_t)ys[r] * sws;
            const WT* wr = SW + (size_t)ys[r] * sws;
            const uint32_t a = gr[xs[0]], b = gr[xs[1]], c = gr[xs[2]], d = gr[xs[3]], e = gr[xs[4]];
            hbr[r] = (c & kM2) * 6u + ((b & kM2) + (d & kM2)) * 4u + (a & kM2) + (e & kM2);
            hg[r] = ((c >> 8) & 0xFFu) * 6u + (((b >> 8) & 0xFFu) + ((d >> 8) & 0xFFu)) * 4u + ((a >> 8) & 0xFFu) + ((e >> 8) & 0xFFu);
            if constexpr (sizeof(WT) == 4) hw[r] = pyr_h(fa, u, (float)wr[xs[0]], (float)wr[xs[1]], (float)wr[xs[2]], (float)wr[xs[3]], (float)wr[xs[4]]);
            else hs[r] = (int)wr[xs[2]] * 6 + ((int)wr[xs[1]] + (int)wr[xs[3]]) * 4 + (int)wr[xs[0]] + (int)wr[xs[4]];
        }
        const uint32_t vbr = hbr[0] + hbr[4] + (hbr[1] + hbr[3]) * 4u + hbr[2] * 6u;
        const uint32_t vg = hg[0] + hg[4] + (hg[1] + hg[3]) * 4u + hg[2] * 6u;
        DG[o] = (((vbr + 0x00800080u) >> 8) & kM2) | (((vg + 128u) >> 8) << 8);
        if constexpr (sizeof(WT) == 4) DW[o] = (WT)pyr_v(fa, u, hw[0], hw[1], hw[2], hw[3], hw[4]);
        else DW[o] = (WT)sat16((hs[0] + hs[4] + (hs[1] + hs[3]) * 4 + hs[2] * 6 + 128) >> 8);   // pyrDown CV_16S
    }
}
cudaError_t launch_rnd_pyrdown(const RenderJob* d_jobs, int n_frames, int max_px_l0, uint8_t* scratch, int level, int f32_mode, int s16_weights,
                               cudaStream_t stream) {
    if (n_frames == 0) return cudaSuccess;
    const int px = std::max(1, max_px_l0 >> (2 * (level + 1)));
    dim3 g((unsigned)std::min((px + 255) / 256, 4096), n_frames);
    if (s16_weights) rnd_pyrdown_kernel<int16_t><<<g, 256, 0, stream>>>(d_jobs, scratch, level, f32_mode);
    else rnd_pyrdown_kernel<float><<<g, 256, 0, stream>>>(d_jobs, scratch, level, f32_mode);
    return cudaGetLastError();
}

// Laplacian of frame J at px (xx, yy) of its level-l sub-image: G_l - pyrUp(G_{l+1}) (createLaplacePyr), or G itself at the top.
__device__ __forceinline__ void rnd_lap_px(const RenderJob& J, const uint8_t* __restrict__ scratch, int l, int levels, int xx, int yy, int lap[3]) {
    const int w = J.sw >> l;
    const uint32_t g = reinterpret_cast<const uint32_t*>(scratch + J.g_off[l])[(size_t)yy * w + xx];
    lap[0] = g & 0xFF; lap[1] = (g >> 8) & 0xFF; lap[2] = (g >> 16) & 0xFF;
    if (l == levels - 1) return;
    const int cw = w >> 1, ch = (J.sh >> l) >> 1;
    const uint32_t* C = reinterpret_cast<const uint32_t*>(scratch + J.g_off[l + 1]);
    const int i = xx >> 1, j = yy >> 1;
    const int c0 = pyrup_axis_lo(i - 1, cw), c2 = pyrup_axis_hi(i + 1, cw);
    const int r0 = pyrup_axis_lo(j - 1, ch), r2 = pyrup_axis_hi(j + 1, ch);
    const bool xodd = xx & 1, yodd = yy & 1;
    const uint32_t *q0 = C + (size_t)r0 * cw, *q1 = C + (size_t)j * cw, *q2 = C + (size_t)r2 * cw;
    uint32_t hbr[3], hg[3];
    const uint32_t* q[3] = {q0, q1, q2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint32_t a = q[k][c0], b = q[k][i], c = q[k][c2];
        if (xodd) { hbr[k] = ((b & kM2) + (c & kM2)) * 4u; hg[k] = (((b >> 8) & 0xFFu) + ((c >> 8) & 0xFFu)) * 4u; }
        else { hbr[k] = (a & kM2) + (b & kM2) * 6u + (c & kM2); hg[k] = ((a >> 8) & 0xFFu) + ((b >> 8) & 0xFFu) * 6u + ((c >> 8) & 0xFFu); }
    }
    uint32_t vbr, vg;
    if (yodd) { vbr = (hbr[1] + hbr[2]) * 4u; vg = (hg[1] + hg[2]) * 4u; }
    else { vbr = hbr[0] + hbr[1] * 6u + hbr[2]; vg = hg[0] + hg[1] * 6u + hg[2]; }
    const uint32_t ubr = ((vbr + 0x00200020u) >> 6) & 0x03FF03FFu;   // <= 255 per lane
    lap[0] -= (int)(ubr & 0xFFFFu);
    lap[1] -= (int)((vg + 32u) >> 6);
    lap[2] -= (int)(ubr >> 16);
}

// ---- canvas-centric blend of a batch (Map2DRender.cpp:199-250) ----
// BLEND 0: `if (w >= dst_w) { dst_w = w; dst = src; }` in feed order == the LAST frame whose weight is >= everything before it.
// BLEND 1: dst += short(src * w) (float product, truncated), dst_w += w, in feed order (float sums are order-dependent).
// BLEND 2: dst += short((src * w) >> 8), dst_w += w on 16-bit lanes (wrapping like the reference's short arithmetic).
template <int BLEND>
__global__ void __launch_bounds__(256) rnd_blend_kernel(const RenderJob* __restrict__ jobs, int n_frames, const uint8_t* __restrict__ scratch,
                                                        const __grid_constant__ RenderCanvas C) {
    extern __shared__ int4 rects[];   // per frame: x_tl, y_tl, sw, sh (level 0)
    for (int k = threadIdx.x; k < n_frames; k += 256) rects[k] = make_int4(jobs[k].x_tl, jobs[k].y_tl, jobs[k].sw, jobs[k].sh);
    __syncthreads();
    long long q = (long long)blockIdx.x * 256 + threadIdx.x;
    int l = 0;
    for (; l < C.levels; l++) {
        const long long cnt = (long long)C.w[l] * C.h[l];
        if (q < cnt) break;
        q -= cnt;
    }
    if (l >= C.levels) return;
    const int y = (int)(q / C.w[l]), x = (int)(q - (long long)y * C.w[l]);
    const size_t o = (size_t)q;
    using WT = typename std::conditional<BLEND == 2, int16_t, float>::type;
    WT* wp = reinterpret_cast<WT*>(C.wgt[l]) + o;
    WT bw = *wp;
    int acc[3] = {0, 0, 0};
    if (BLEND != 0) { acc[0] = C.lap[l][0][o]; acc[1] = C.lap[l][1][o]; acc[2] = C.lap[l][2][o]; }
    int best = -1;
    bool touched = false;
    for (int k = 0; k < n_frames; k++) {
        const int4 r = rects[k];
        const int xx = x - (r.x >> l), yy = y - (r.y >> l), w = r.z >> l, h = r.w >> l;
        if ((unsigned)xx >= (unsigned)w || (unsigned)yy >= (unsigned)h) continue;
        const RenderJob& J = jobs[k];
        const WT wv = reinterpret_cast<const WT*>(scratch + J.w_off[l])[(size_t)yy * w + xx];
        if (BLEND == 0) {
            if (wv >= bw) { bw = wv; best = k; }
        } else {
            int lap[3];
            rnd_lap_px(J, scratch, l, C.levels, xx, yy, lap);
            touched = true;
            if (BLEND == 1) {
#pragma unroll
                for (int c = 0; c < 3; c++) acc[c] = (int)(int16_t)(acc[c] + (int)(int16_t)(int)((float)lap[c] * (float)wv));
                bw = (WT)((float)bw + (float)wv);
            } else {
#pragma unroll
                for (int c = 0; c < 3; c++) acc[c] = (int)(int16_t)(acc[c] + (int)(int16_t)((lap[c] * (int)wv) >> 8));
                bw = (WT)(int16_t)((int)bw + (int)wv);
            }
        }
    }
    if (BLEND == 0) {
        if (best < 0) return;
        int lap[3];
        const int4 r = rects[best];
        rnd_lap_px(jobs[best], scratch, l, C.levels, x - (r.x >> l), y - (r.y >> l), lap);
        acc[0] = lap[0]; acc[1] = lap[1]; acc[2] = lap[2];
    } else if (!touched) return;
    *wp = bw;
    C.lap[l][0][o] = (int16_t)acc[0]; C.lap[l][1][o] = (int16_t)acc[1]; C.lap[l][2][o] = (int16_t)acc[2];
}
static long long canvas_px(const RenderCanvas& C) {
    long long n = 0;
    for (int l = 0; l < C.levels; l++) n += (long long)C.w[l] * C.h[l];
    return n;
}
cudaError_t launch_rnd_blend(const RenderJob* d_jobs, int n_frames, const uint8_t* scratch, const RenderCanvas& C, int blend, cudaStream_t stream) {
    if (n_frames == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((canvas_px(C) + 255) / 256);
    const size_t smem = (size_t)n_frames * sizeof(int4);
    if (blend == 0) rnd_blend_kernel<0><<<blocks, 256, smem, stream>>>(d_jobs, n_frames, scratch, C);
    else if (blend == 1) rnd_blend_kernel<1><<<blocks, 256, smem, stream>>>(d_jobs, n_frames, scratch, C);
    else rnd_blend_kernel<2><<<blocks, 256, smem, stream>>>(d_jobs, n_frames, scratch, C);
    return cudaGetLastError();
}

// ---- normalizeUsingWeightMap (cv::detail, stitching/blenders.cpp), all levels in one launch ----
template <int BLEND>
__global__ void __launch_bounds__(256) rnd_normalize_kernel(const __grid_constant__ RenderCanvas C) {
    long long q = (long long)blockIdx.x * 256 + threadIdx.x;
    int l = 0;
    for (; l < C.levels; l++) {
        const long long cnt = (long long)C.w[l] * C.h[l];
        if (q < cnt) break;
        q -= cnt;
    }
    if (l >= C.levels) return;
    const size_t o = (size_t)q;
    if (BLEND == 1) {
        const float d = reinterpret_cast<const float*>(C.wgt[l])[o] + 1e-5f;
#pragma unroll
        for (int c = 0; c < 3; c++) C.lap[l][c][o] = (int16_t)(int)((float)C.lap[l][c][o] / d);
    } else {
        const int w = (int)reinterpret_cast<const int16_t*>(C.wgt[l])[o] + 1;
#pragma unroll
        for (int c = 0; c < 3; c++) C.lap[l][c][o] = (int16_t)(w ? ((int)C.lap[l][c][o] * 256) / w : 0);
    }
}
cudaError_t launch_rnd_normalize(const RenderCanvas& C, int blend, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((canvas_px(C) + 255) / 256);
    if (blend == 1) rnd_normalize_kernel<1><<<blocks, 256, 0, stream>>>(C);
    else if (blend == 2) rnd_normalize_kernel<2><<<blocks, 256, 0, stream>>>(C);
    return cudaGetLastError();
}

// ---- Blender::blend: dst_mask = weight > WEIGHT_EPS, dst.setTo(0, mask == 0), crop to dst_roi_final_; then 16S -> 8U ----
__global__ void __launch_bounds__(256) rnd_final_kernel(const __grid_constant__ RenderCanvas C, int s16_weights, int wf, int hf, int16_t* __restrict__ out16,
                                                        uint8_t* __restrict__ out8, uint8_t* __restrict__ mask) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)wf * hf) return;
    const int y = (int)(i / wf), x = (int)(i - (long long)y * wf);
    const size_t s = (size_t)y * C.w[0] + x;
    const bool on = s16_weights ? reinterpret_cast<const int16_t*>(C.wgt[0])[s] > 0 : reinterpret_cast<const float*>(C.wgt[0])[s] > (float)1e-5;
    mask[i] = on ? 255 : 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int v = on ? (int)C.lap[0][c][s] : 0;
        out16[3 * i + c] = (int16_t)v;
        out8[3 * i + c] = (uint8_t)min(max(v, 0), 255);
    }
}
cudaError_t launch_rnd_final(const RenderCanvas& C, int s16_weights, int wf, int hf, int16_t* out16, uint8_t* out8, uint8_t* mask, cudaStream_t stream) {
    const unsigned blocks = (unsigned)(((long long)wf * hf + 255) / 256);
    rnd_final_kernel<<<blocks, 256, 0, stream>>>(C, s16_weights, wf, hf, out16, out8, mask);
    return cudaGetLastError();
}

}  // namespace m2d
