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
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int sat16(int v) { return min(max(v, -32768), 32767); }

// cv::warpPerspectiveInvoker coordinate for destination px (x,y): the row base is formed at the first column of
// the 64-px block and advanced by M0*x1 (same association as OpenCV; region widths are multiples of 256).
// Returns the un-quantised source coordinate; INTER_LINEAR uses rint(32*f), INTER_NEAREST rint(f) — 32/W is
// exactly 32*(1/W) in binary floating point, so one division serves both.
__device__ __forceinline__ void warp_coord(const double* M, int x, int y, double& fx, double& fy) {
    int xb = x & ~63, x1 = x & 63;
    double X0 = M[0] * xb + M[1] * y + M[2];
    double Y0 = M[3] * xb + M[4] * y + M[5];
    double W0 = M[6] * xb + M[7] * y + M[8];
    double W = W0 + M[6] * x1;
    W = (W != 0.0) ? 1.0 / W : 0.0;
    fx = (X0 + M[0] * x1) * W;
    fy = (Y0 + M[3] * x1) * W;
}
__device__ __forceinline__ int round_coord(double f) {  // saturate_cast<int>(max(INT_MIN, min(INT_MAX, f)))
    return __double2int_rn(fmax(-2147483648.0, fmin(2147483647.0, f)));
}

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
// weighted mode: fused warp (8UC4 bilinear, constant-0 border) + max-alpha select  (Map2DCPU.cpp:282-333)
// One CTA = 4 rows x 256 px of one tile; one thread = 4 consecutive px (one 16-byte state vector).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fetch_bgra(const WeightedParams& p, int sx, int sy) {
    if ((unsigned)sx >= (unsigned)p.sw || (unsigned)sy >= (unsigned)p.sh) return 0u;
    const uint8_t* q = p.src + (size_t)sy * p.src_stride + 3 * sx;
    uint32_t a = __ldg(p.alpha + (size_t)sy * p.sw + sx);
    return (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16) | (a << 24);
}

__device__ __forceinline__ uint32_t sample_bgra(const WeightedParams& p, int x, int y) {
    double fx, fy;
    warp_coord(p.hinv, x, y, fx, fy);
    int X = round_coord(fx * 32.0), Y = round_coord(fy * 32.0);
    int sx = clampi(X >> 5, -32768, 32767), sy = clampi(Y >> 5, -32768, 32767);
    if (sx >= p.sw || sx + 1 < 0 || sy >= p.sh || sy + 1 < 0) return 0u;
    int a = X & 31, b = Y & 31;
    uint32_t w00 = (32 - a) * (32 - b), w01 = a * (32 - b), w10 = (32 - a) * b, w11 = a * b;
    uint32_t v00 = fetch_bgra(p, sx, sy), v01 = fetch_bgra(p, sx + 1, sy);
    uint32_t v10 = fetch_bgra(p, sx, sy + 1), v11 = fetch_bgra(p, sx + 1, sy + 1);
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int sh = 8 * c;
        uint32_t v = ((v00 >> sh) & 255u) * w00 + ((v01 >> sh) & 255u) * w01 + ((v10 >> sh) & 255u) * w10 +
                     ((v11 >> sh) & 255u) * w11;
        out |= ((v + 512u) >> 10) << sh;  // == (sum*32 + 16384) >> 15 of FixedPtCast<int,uchar,15>; <= 255
    }
    return out;
}

__global__ void __launch_bounds__(256) weighted_fuse_kernel(const __grid_constant__ WeightedParams p) {
    int twx = blockIdx.x % p.r.wnx, twy = blockIdx.x / p.r.wnx;
    int tx = p.r.wx0 + twx, ty = p.r.wy0 + twy;
    uint8_t* tp = p.table[(size_t)ty * p.grid_w + tx];
    if (!tp) return;
    int rtx = tx - p.r.rx0, rty = ty - p.r.ry0;
    int bit = rty * p.r.nx + rtx;
    bool fresh = (p.fresh[bit >> 5] >> (bit & 31)) & 1u;
    int px = (threadIdx.x & 63) * 4, py = blockIdx.y * 4 + (threadIdx.x >> 6);
    int X = rtx * kEle + px, Y = rty * kEle + py;
    uint4* sp = reinterpret_cast<uint4*>(tp + ((size_t)py * kEle + px) * 4);
    uint4 st = fresh ? make_uint4(0u, 0u, 0u, 0u) : *sp;
    uint32_t s[4] = {st.x, st.y, st.z, st.w};
    bool changed = fresh;
    unsigned wins = 0, foot = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t d = sample_bgra(p, X + i, Y);
        foot += (d >> 24) != 0u;
        if ((s[i] >> 24) < (d >> 24)) {  // strict '<' : Map2DCPU.cpp:327
            s[i] = d;
            changed = true;
            wins++;
        }
    }
    if (changed) *sp = make_uint4(s[0], s[1], s[2], s[3]);
    if (p.stats) {
        unsigned long long f = warp_sum(foot), w = warp_sum(fresh ? 0u : wins);
        if ((threadIdx.x & 31) == 0) {
            if (f) atomicAdd(p.stats + 0, f);
            if (w) atomicAdd(p.stats + 1, w);
        }
    }
}
cudaError_t launch_weighted(const WeightedParams& p, cudaStream_t stream) {
    dim3 g(p.r.wnx * p.r.wny, kEle / 4);
    weighted_fuse_kernel<<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// multi-band, stage 1: warp the frame into level 0 of the scratch pyramid over the window
//   image : 16SC3 bilinear with BORDER_REFLECT, exact integer form of remapBilinear<Cast<float,short>> + cvRound
//   weight: nearest from the float weight image, constant-0 border
// One thread = 2 horizontally adjacent px; planar outputs.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mb_sample(const MultibandParams& p, int x, int y, int& b, int& g, int& r, float& w) {
    double fx, fy;
    warp_coord(p.hinv, x, y, fx, fy);
    // weight: INTER_NEAREST
    int nx = clampi(round_coord(fx), -32768, 32767), ny = clampi(round_coord(fy), -32768, 32767);
    w = ((unsigned)nx < (unsigned)p.sw && (unsigned)ny < (unsigned)p.sh) ? __ldg(p.wimg + (size_t)ny * p.sw + nx) : 0.f;
    // image: INTER_LINEAR + BORDER_REFLECT
    int X = round_coord(fx * 32.0), Y = round_coord(fy * 32.0);
    int sx = clampi(X >> 5, -32768, 32767), sy = clampi(Y >> 5, -32768, 32767);
    int a = X & 31, bb = Y & 31;
    int w00 = (32 - a) * (32 - bb), w01 = a * (32 - bb), w10 = (32 - a) * bb, w11 = a * bb;
    int sx0 = reflect_idx(sx, p.sw), sx1 = reflect_idx(sx + 1, p.sw);
    int sy0 = reflect_idx(sy, p.sh), sy1 = reflect_idx(sy + 1, p.sh);
    const uint8_t* r0 = p.src + (size_t)sy0 * p.src_stride;
    const uint8_t* r1 = p.src + (size_t)sy1 * p.src_stride;
    const uint8_t *p00 = r0 + 3 * sx0, *p01 = r0 + 3 * sx1, *p10 = r1 + 3 * sx0, *p11 = r1 + 3 * sx1;
    int out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // the float sum S00*w0+S01*w1+S10*w2+S11*w3 is exact (<= 18 bits), so cvRound(sum) == RNE(v/1024)
        int v = (int)__ldg(p00 + c) * w00 + (int)__ldg(p01 + c) * w01 + (int)__ldg(p10 + c) * w10 + (int)__ldg(p11 + c) * w11;
        int q = v >> 10, rem = v & 1023;
        q += (rem > 512) || (rem == 512 && (q & 1));
        out[c] = q;
    }
    b = out[0]; g = out[1]; r = out[2];
}

__global__ void __launch_bounds__(256) mb_warp_kernel(const __grid_constant__ MultibandParams p) {
    const PyrLevel& L = p.lv[0];
    int u = (blockIdx.x * 32 + threadIdx.x) * 2, v = blockIdx.y * 8 + threadIdx.y;
    if (u >= L.ww || v >= L.wh) return;
    int x = u + L.ox, y = v + L.oy;
    int b0, g0, r0, b1, g1, r1;
    float w0, w1;
    mb_sample(p, x, y, b0, g0, r0, w0);
    mb_sample(p, x + 1, y, b1, g1, r1, w1);
    size_t o = (size_t)v * L.ww + u;
    *reinterpret_cast<short2*>(L.g[0] + o) = make_short2((short)b0, (short)b1);
    *reinterpret_cast<short2*>(L.g[1] + o) = make_short2((short)g0, (short)g1);
    *reinterpret_cast<short2*>(L.g[2] + o) = make_short2((short)r0, (short)r1);
    *reinterpret_cast<float2*>(L.w + o) = make_float2(w0, w1);
}
cudaError_t launch_mb_warp(const MultibandParams& p, cudaStream_t stream) {
    dim3 b(32, 8), g((p.lv[0].ww / 2 + 31) / 32, (p.lv[0].wh + 7) / 8);
    mb_warp_kernel<<<g, b, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// multi-band, stage 2: pyrDown level l -> l+1 of the scratch pyramid (3 int16 planes + f32 weight plane).
// Border = BORDER_REFLECT_101 in REGION coordinates (the window may be a sub-rectangle when tiles are sharded).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mb_pyrdown_kernel(const __grid_constant__ MultibandParams p, int l) {
    const PyrLevel& S = p.lv[l];
    const PyrLevel& D = p.lv[l + 1];
    int u = blockIdx.x * 32 + threadIdx.x, v = blockIdx.y * 8 + threadIdx.y;
    if (u >= D.ww || v >= D.wh) return;
    int U = u + D.ox, V = v + D.oy;
    int xs[5], ys[5];
#pragma unroll
    for (int d = 0; d < 5; d++) {
        xs[d] = clampi(reflect101_idx(2 * U + d - 2, S.rw) - S.ox, 0, S.ww - 1);
        ys[d] = clampi(reflect101_idx(2 * V + d - 2, S.rh) - S.oy, 0, S.wh - 1);
    }
    size_t o = (size_t)v * D.ww + u;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int16_t* G = S.g[c];
        int acc = 0;
#pragma unroll
        for (int d = 0; d < 5; d++) {
            const int16_t* row = G + (size_t)ys[d] * S.ww;
            int h = row[xs[2]] * 6 + (row[xs[1]] + row[xs[3]]) * 4 + row[xs[0]] + row[xs[4]];
            const int kv = (d == 0 || d == 4) ? 1 : ((d == 2) ? 6 : 4);
            acc += kv * h;
        }
        D.g[c][o] = (int16_t)sat16((acc + 128) >> 8);
    }
    {
        // f32, OpenCV 2.4.9 association: rows s0*6 + (s-1+s1)*4 + s-2 + s2 (left to right);
        // columns ((r0+r4)+(r2+r2)) + ((r1+r3)+r2)*4, scaled by 1/256.
        float h[5];
#pragma unroll
        for (int d = 0; d < 5; d++) {
            const float* row = S.w + (size_t)ys[d] * S.ww;
            h[d] = row[xs[2]] * 6.f + (row[xs[1]] + row[xs[3]]) * 4.f + row[xs[0]] + row[xs[4]];
        }
        float t0 = (h[0] + h[4]) + (h[2] + h[2]);
        float t1 = (h[1] + h[3]) + h[2];
        D.w[o] = (t0 + t1 * 4.f) * (1.f / 256.f);
    }
}
cudaError_t launch_mb_pyrdown(const MultibandParams& p, int level, cudaStream_t stream) {
    const PyrLevel& D = p.lv[level + 1];
    dim3 b(32, 8), g((D.ww + 31) / 32, (D.wh + 7) / 8);
    mb_pyrdown_kernel<<<g, b, 0, stream>>>(p, level);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// multi-band, stage 3: Laplacian (G_l - pyrUp(G_{l+1}), never materialised) + per-band '>=' select into the
// tile state, all levels in one launch.  One thread = 2 horizontally adjacent px of one level of one tile.
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

__global__ void __launch_bounds__(256) mb_select_kernel(const __grid_constant__ MultibandParams p,
                                                        const __grid_constant__ TileLayout lay) {
    int twx = blockIdx.x % p.r.wnx, twy = blockIdx.x / p.r.wnx;
    int tx = p.r.wx0 + twx, ty = p.r.wy0 + twy;
    uint8_t* tp = p.table[(size_t)ty * p.grid_w + tx];
    if (!tp) return;
    int rtx = tx - p.r.rx0, rty = ty - p.r.ry0;
    int bit = rty * p.r.nx + rtx;
    bool fresh = (p.fresh[bit >> 5] >> (bit & 31)) & 1u;

    // flat pair index -> (level, row, column pair)
    int pair = blockIdx.y * 256 + threadIdx.x;
    int l = 0, n = kEle, half = kEle / 2;
    for (; l < p.levels; l++) {
        n = kEle >> l;
        half = n > 1 ? n / 2 : 1;
        int cnt = n * half;
        if (pair < cnt) break;
        pair -= cnt;
    }
    bool valid = l < p.levels;
    if (!valid) { l = p.levels - 1; n = kEle >> l; half = n > 1 ? n / 2 : 1; pair = 0; }
    int py = pair / half, px = (pair % half) * 2;
    bool two = n > 1;
    const PyrLevel& L = p.lv[l];
    int X = rtx * n + px, Y = rty * n + py;  // region coordinates at level l
    size_t so = (size_t)(Y - L.oy) * L.ww + (X - L.ox);
    float sw0 = L.w[so], sw1 = two ? L.w[so + 1] : 0.f;
    size_t to = (size_t)py * n + px;
    float* tw = reinterpret_cast<float*>(tp + lay.wgt_off[l]) + to;
    bool win0, win1;
    if (fresh) { win0 = true; win1 = two; }
    else {
        float dw0 = tw[0], dw1 = two ? tw[1] : 0.f;
        win0 = sw0 >= dw0;            // '>=' : MultiBandMap2DCPU.cpp:542 (0 >= 0 ties overwrite)
        win1 = two && (sw1 >= dw1);
    }
    win0 = win0 && valid; win1 = win1 && valid;
    if (p.stats) {  // block-uniform branch: every lane reaches the shuffle
        unsigned long long w = warp_sum(fresh ? 0u : ((unsigned)win0 + (unsigned)win1));
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(p.stats + l, w);
    }
    if (!win0 && !win1) return;

    int lap0[3], lap1[3];
    if (l == p.levels - 1) {
#pragma unroll
        for (int c = 0; c < 3; c++) { lap0[c] = L.g[c][so]; lap1[c] = two ? L.g[c][so + 1] : 0; }
    } else {
        const PyrLevel& C = p.lv[l + 1];
        int i = X >> 1, j = Y >> 1;
        int c0 = clampi(pyrup_axis_lo(i - 1, C.rw) - C.ox, 0, C.ww - 1);
        int c1 = clampi(i - C.ox, 0, C.ww - 1);
        int c2 = clampi(pyrup_axis_hi(i + 1, C.rw) - C.ox, 0, C.ww - 1);
        int r0 = clampi(pyrup_axis_lo(j - 1, C.rh) - C.oy, 0, C.wh - 1);
        int r1 = clampi(j - C.oy, 0, C.wh - 1);
        int r2 = clampi(pyrup_axis_hi(j + 1, C.rh) - C.oy, 0, C.wh - 1);
        bool yodd = Y & 1;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int16_t* G = C.g[c];
            const int16_t *q0 = G + (size_t)r0 * C.ww, *q1 = G + (size_t)r1 * C.ww, *q2 = G + (size_t)r2 * C.ww;
            int e0 = q0[c0] + q0[c1] * 6 + q0[c2], o0 = (q0[c1] + q0[c2]) * 4;   // even / odd column, row j-1
            int e1 = q1[c0] + q1[c1] * 6 + q1[c2], o1 = (q1[c1] + q1[c2]) * 4;   // row j
            int e2 = q2[c0] + q2[c1] * 6 + q2[c2], o2 = (q2[c1] + q2[c2]) * 4;   // row j+1
            int ve = yodd ? (e1 + e2) * 4 : (e0 + e1 * 6 + e2);
            int vo = yodd ? (o1 + o2) * 4 : (o0 + o1 * 6 + o2);
            int upe = sat16((ve + 32) >> 6), upo = sat16((vo + 32) >> 6);
            lap0[c] = sat16((int)L.g[c][so] - upe);
            lap1[c] = two ? sat16((int)L.g[c][so + 1] - upo) : 0;
        }
    }
    size_t plane = (size_t)n * n;
    int16_t* tl = reinterpret_cast<int16_t*>(tp + lay.lap_off[l]) + to;
    if (!two) {
        tl[0] = (int16_t)lap0[0]; tl[plane] = (int16_t)lap0[1]; tl[2 * plane] = (int16_t)lap0[2];
        tw[0] = sw0;
        return;
    }
    if (win0 && win1) {
#pragma unroll
        for (int c = 0; c < 3; c++)
            *reinterpret_cast<short2*>(tl + c * plane) = make_short2((short)lap0[c], (short)lap1[c]);
        *reinterpret_cast<float2*>(tw) = make_float2(sw0, sw1);
    } else if (win0) {
#pragma unroll
        for (int c = 0; c < 3; c++) tl[c * plane] = (int16_t)lap0[c];
        tw[0] = sw0;
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) tl[c * plane + 1] = (int16_t)lap1[c];
        tw[1] = sw1;
    }
}
cudaError_t launch_mb_select(const MultibandParams& p, const TileLayout& lay, cudaStream_t stream) {
    int pairs = 0;
    for (int l = 0; l < p.levels; l++) {
        int n = kEle >> l;
        pairs += n * (n > 1 ? n / 2 : 1);
    }
    dim3 g(p.r.wnx * p.r.wny, (pairs + 255) / 256);
    mb_select_kernel<<<g, 256, 0, stream>>>(p, lay);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// collapse (MultiBandMap2DCPU::save, :779-841): paste tiles into per-level mosaics, restore from the Laplacian
// pyramid coarse -> fine (pyrUp + saturating add), convert to 8-bit and paint the background where weight == 0.
// ---------------------------------------------------------------------------------------------------------
__global__ void mosaic_clear_kernel(MosaicLevel m, float* w0) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)m.w * m.h;
    if (i >= n) return;
    m.g[0][i] = 0; m.g[1][i] = 0; m.g[2][i] = 0;
    if (w0) w0[i] = 0.f;
}
cudaError_t launch_mosaic_clear(MosaicLevel m, float* w0, cudaStream_t stream) {
    size_t n = (size_t)m.w * m.h;
    mosaic_clear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(m, w0);
    return cudaGetLastError();
}
__global__ void mosaic_paste_kernel(const uint8_t* __restrict__ tile, size_t lap_off, size_t wgt_off, int n, MosaicLevel m,
                                    float* w0, int tx, int ty) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    int y = i / n, x = i % n;
    const int16_t* tl = reinterpret_cast<const int16_t*>(tile + lap_off);
    size_t o = (size_t)(ty * n + y) * m.w + (size_t)tx * n + x;
    size_t plane = (size_t)n * n;
    m.g[0][o] = tl[i]; m.g[1][o] = tl[plane + i]; m.g[2][o] = tl[2 * plane + i];
    if (w0) w0[o] = reinterpret_cast<const float*>(tile + wgt_off)[i];
}
cudaError_t launch_mosaic_paste(const uint8_t* tile, const TileLayout& lay, int level, MosaicLevel m, float* w0, int tx,
                                int ty, cudaStream_t stream) {
    int n = kEle >> level;
    mosaic_paste_kernel<<<(n * n + 255) / 256, 256, 0, stream>>>(tile, lay.lap_off[level], lay.wgt_off[level], n, m,
                                                                 level == 0 ? w0 : nullptr, tx, ty);
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
