// kernels.cu — hand-written sm_100a kernels of the Map2D feed() path.
//
// All arithmetic follows the exact integer / float recipes of the OpenCV primitives the reference calls
// (SURVEY.md §9): 1/32-px coordinate quantisation with round-half-even, fixed-point bilinear for 8UC4, exact
// bilinear + cvRound for 16SC3 with BORDER_REFLECT, nearest for the float weight, [1 4 6 4 1] pyrDown with
// (x+128)>>8, pyrUp with (x+32)>>6 and its asymmetric border.  Compiled with --fmad=false: no float or double
// contraction anywhere, so results match the CPU oracle bit for bit.
#include "kernels.cuh"
#include "device_common.cuh"

namespace m2d {


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
// multi-band stage 2: pyrDown level l -> l+1 for every frame of the group (u8x4 Gaussian + f32 weight).
// Separable [1 4 6 4 1], BORDER_REFLECT_101 in REGION coordinates.  The int path runs on packed 16-bit lanes
// (B,R in one register, G alone): row sums <= 4080, column sums <= 65280, (x+128)>>8 -- no lane ever overflows.
// One thread = 2 adjacent output columns x 4 output rows: 11 input rows stream through a 5-row register window,
// each fetched with four 8-byte loads per plane (7 input columns), so every input px is loaded ~1.1x.
// ---------------------------------------------------------------------------------------------------------
struct HRow { uint32_t br0, g0, br1, g1; float w0, w1; };  // horizontal sums for output columns u and u+1

__device__ __forceinline__ HRow pyr_hrow(const uint32_t* __restrict__ gr, const float* __restrict__ wr, const int* xs, bool fast, const F32Assoc& fa, int U) {
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
    h.w0 = pyr_h(fa, U, f[0], f[1], f[2], f[3], f[4]);
    h.w1 = pyr_h(fa, U + 1, f[2], f[3], f[4], f[5], f[6]);
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
    const F32Assoc fa = f32_assoc(p.f32_mode, srw);
    HRow h[5];
#pragma unroll
    for (int r = 0; r < 11; r++) {
        int ys = clampi(reflect101_idx(2 * V0 + r - 2, srh) - soy, 0, swh - 1);
        h[r % 5] = pyr_hrow(SG + (size_t)ys * sww, SW + (size_t)ys * sww, xs, fast, fa, U);
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
                float ow0 = pyr_v(fa, U, r0.w0, r1.w0, r2.w0, r3.w0, r4.w0), ow1 = pyr_v(fa, U + 1, r0.w1, r1.w1, r2.w1, r3.w1, r4.w1);
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
        const F32Assoc fa = f32_assoc(p.f32_mode, srw);
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
                hw[r] = pyr_h(fa, U, wr[xs[0]], wr[xs[1]], wr[xs[2]], wr[xs[3]], wr[xs[4]]);
            }
            uint32_t vbr = hbr[0] + hbr[4] + (hbr[1] + hbr[3]) * 4u + hbr[2] * 6u;
            uint32_t vg = hg[0] + hg[4] + (hg[1] + hg[3]) * 4u + hg[2] * 6u;
            DG[o] = (((vbr + 0x00800080u) >> 8) & kM2) | (((vg + 128u) >> 8) << 8);
            DW[o] = pyr_v(fa, U, hw[0], hw[1], hw[2], hw[3], hw[4]);
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
    t.cmin_off = off;                       // per-(level, cell) lower bound of the weight planes, see kernels.cuh
    off += (size_t)levels * 64 * sizeof(float);
    t.bytes = (off + 255) & ~(size_t)255;
    return t;
}


__global__ void __launch_bounds__(256, 6) mb_select_kernel(const __grid_constant__ GroupParams p, const __grid_constant__ TileLayout lay) {
    const TileWork T = p.tiles[blockIdx.x];
    // cmin (the per-cell lower bound of the tile's weights that the weights-first pipeline culls against) only has to be a
    // LOWER bound, and weights never decrease: this pipeline leaves it as it is, and starts a fresh tile at 0
    if (T.fresh && blockIdx.y == 0) {
        float* cmin = reinterpret_cast<float*>(T.state + lay.cmin_off);
        for (int i = threadIdx.x; i < lay.levels * 64; i += 256) cmin[i] = 0.f;
    }
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

// One launch pastes every touched tile into the zero-initialised per-level mosaics: levels whose tile side is >= 8 px move 8 px
// (16 bytes per plane) per thread -- int16 planes moved two bytes per thread use a sixth of what a warp's load / store path can
// carry -- and the few px of deeper levels go through a one-warp tail kernel.
__global__ void __launch_bounds__(256) mosaic_paste8_kernel(const PasteItem* __restrict__ items, const __grid_constant__ TileLayout lay,
                                                            const __grid_constant__ MosaicSet ms, int vec_levels) {
    const PasteItem it = items[blockIdx.x];
    int u = blockIdx.y * 256 + threadIdx.x;      // 8-px unit index over the vector levels
    int l = 0;
    for (; l < vec_levels; l++) {
        const int cnt = (lay.px_off[l + 1] - lay.px_off[l]) >> 3;
        if (u < cnt) break;
        u -= cnt;
    }
    if (l >= vec_levels) return;
    const int n = kEle >> l, upr = n >> 3;
    const int y = u / upr, x = (u - y * upr) * 8;
    const int16_t* tl = reinterpret_cast<const int16_t*>(it.tile + lay.lap_off[l]);
    const MosaicLevel& m = ms.lv[l];
    const size_t s = (size_t)y * n + x, plane = (size_t)n * n;
    const size_t o = (size_t)(it.ty * n + y) * m.w + (size_t)it.tx * n + x;
#pragma unroll
    for (int c = 0; c < 3; c++) *reinterpret_cast<uint4*>(m.g[c] + o) = *reinterpret_cast<const uint4*>(tl + c * plane + s);
    if (l == 0) {
        const float4* w = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(it.tile + lay.wgt_off[0]) + s);
        float4* d = reinterpret_cast<float4*>(ms.w0 + o);
        d[0] = w[0]; d[1] = w[1];
    }
}
__global__ void __launch_bounds__(256) mosaic_paste_tail_kernel(const PasteItem* __restrict__ items, const __grid_constant__ TileLayout lay,
                                                                const __grid_constant__ MosaicSet ms, int first_level) {
    const PasteItem it = items[blockIdx.x];
    int i = lay.px_off[first_level] + threadIdx.x;   // the levels below 8 px a side hold at most 16 + 4 + 1 px
    if (i >= lay.px_off[lay.levels]) return;
    int l = first_level;
    while (i >= lay.px_off[l + 1]) l++;
    i -= lay.px_off[l];
    const int n = kEle >> l;
    int y = i / n, x = i - y * n;
    const int16_t* tl = reinterpret_cast<const int16_t*>(it.tile + lay.lap_off[l]);
    const MosaicLevel& m = ms.lv[l];
    size_t o = (size_t)(it.ty * n + y) * m.w + (size_t)it.tx * n + x;
    size_t plane = (size_t)n * n;
    m.g[0][o] = tl[i]; m.g[1][o] = tl[plane + i]; m.g[2][o] = tl[2 * plane + i];
}
cudaError_t launch_mosaic_paste(const PasteItem* d_items, int n_items, const TileLayout& lay, const MosaicSet& ms, cudaStream_t stream) {
    if (n_items == 0) return cudaSuccess;
    int vec_levels = 0;
    while (vec_levels < lay.levels && (kEle >> vec_levels) >= 8) vec_levels++;
    if (vec_levels > 0) {
        dim3 g(n_items, ((lay.px_off[vec_levels] >> 3) + 255) / 256);
        mosaic_paste8_kernel<<<g, 256, 0, stream>>>(d_items, lay, ms, vec_levels);
    }
    if (vec_levels < lay.levels) mosaic_paste_tail_kernel<<<n_items, 32, 0, stream>>>(d_items, lay, ms, vec_levels);
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
// The same, 8 consecutive fine px (16 bytes per plane) per thread: needs F.w % 8 == 0 (then C.w == F.w / 2 is a multiple of 4 and
// every row of both levels starts 8-byte aligned).  The 8 px use coarse columns i0-1 .. i0+4 (i0 = x0 / 2) of rows r0, j, r2.
__global__ void __launch_bounds__(256) mosaic_upadd8_kernel(MosaicLevel C, MosaicLevel F) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8, y = blockIdx.y * 8 + threadIdx.y;
    if (x0 >= F.w || y >= F.h) return;
    const int i0 = x0 >> 1, j = y >> 1;
    const int cl = pyrup_axis_lo(i0 - 1, C.w), cr = pyrup_axis_hi(i0 + 4, C.w);
    const int r0 = pyrup_axis_lo(j - 1, C.h), r2 = pyrup_axis_hi(j + 1, C.h);
    const bool yodd = y & 1;
    const size_t o = (size_t)y * F.w + x0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int16_t* G = C.g[c];
        int h[3][8];
        const int rr[3] = {r0, j, r2};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (k == 0 && yodd) continue;        // odd rows use rows j and r2 only
            const int16_t* q = G + (size_t)rr[k] * C.w;
            const short4 mid = *reinterpret_cast<const short4*>(q + i0);
            const int v[6] = {(int)q[cl], (int)mid.x, (int)mid.y, (int)mid.z, (int)mid.w, (int)q[cr]};
#pragma unroll
            for (int m = 0; m < 4; m++) {
                h[k][2 * m] = v[m] + v[m + 1] * 6 + v[m + 2];
                h[k][2 * m + 1] = (v[m + 1] + v[m + 2]) * 4;
            }
        }
        int16_t* fp = F.g[c] + o;
        uint4 fv = *reinterpret_cast<const uint4*>(fp);
        int16_t* f = reinterpret_cast<int16_t*>(&fv);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int v = yodd ? (h[1][e] + h[2][e]) * 4 : (h[0][e] + h[1][e] * 6 + h[2][e]);
            f[e] = (int16_t)sat16((int)f[e] + sat16((v + 32) >> 6));
        }
        *reinterpret_cast<uint4*>(fp) = fv;
    }
}
cudaError_t launch_mosaic_upadd(MosaicLevel coarse, MosaicLevel fine, cudaStream_t stream) {
    dim3 b(32, 8);
    if (fine.w % 8 == 0 && coarse.w * 2 == fine.w) {
        dim3 g((fine.w / 8 + 31) / 32, (fine.h + 7) / 8);
        mosaic_upadd8_kernel<<<g, b, 0, stream>>>(coarse, fine);
    } else {
        dim3 g((fine.w + 31) / 32, (fine.h + 7) / 8);
        mosaic_upadd_kernel<<<g, b, 0, stream>>>(coarse, fine);
    }
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
// 4 px per thread (the mosaic is whole tiles wide): 8-byte loads per plane, one 16-byte weight load, three 4-byte stores.
__global__ void __launch_bounds__(256) mosaic_final4_kernel(MosaicLevel m, const float* __restrict__ w0, int background, uint8_t* __restrict__ out) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n4 = ((size_t)m.w * m.h) >> 2;
    if (q >= n4) return;
    const size_t i = q * 4;
    const float4 w = *reinterpret_cast<const float4*>(w0 + i);
    const short4 b = *reinterpret_cast<const short4*>(m.g[0] + i), g = *reinterpret_cast<const short4*>(m.g[1] + i), r = *reinterpret_cast<const short4*>(m.g[2] + i);
    const uint32_t bg = (uint32_t)min(max(background, 0), 255);
    const float ws[4] = {w.x, w.y, w.z, w.w};
    const int bs[4] = {b.x, b.y, b.z, b.w}, gs[4] = {g.x, g.y, g.z, g.w}, rs[4] = {r.x, r.y, r.z, r.w};
    uint32_t px[12];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const bool off = ws[e] == 0.f;
        px[3 * e] = off ? bg : (uint32_t)min(max(bs[e], 0), 255);
        px[3 * e + 1] = off ? bg : (uint32_t)min(max(gs[e], 0), 255);
        px[3 * e + 2] = off ? bg : (uint32_t)min(max(rs[e], 0), 255);
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(out + 3 * i);
    o[0] = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
    o[1] = px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24);
    o[2] = px[8] | (px[9] << 8) | (px[10] << 16) | (px[11] << 24);
}
cudaError_t launch_mosaic_final(MosaicLevel m0, const float* w0, int background, uint8_t* out_bgr, cudaStream_t stream) {
    size_t n = (size_t)m0.w * m0.h;
    if (n % 4 == 0 && (reinterpret_cast<uintptr_t>(out_bgr) & 3) == 0) mosaic_final4_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, stream>>>(m0, w0, background, out_bgr);
    else mosaic_final_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(m0, w0, background, out_bgr);
    return cudaGetLastError();
}

}  // namespace m2d