// bounds.h — closed-form lower / upper bounds of a frame's multi-band WEIGHT over one cell of one pyramid level.
//
// The weights-first pipeline (kernels_wf.cu) uses them to decide, before any weight is computed, in which cells a
// frame can possibly win a px ("competitive" cells): a frame whose upper bound lies below the best lower bound among
// the other covering frames (or the tile state) loses everywhere in the cell, strictly, and is never evaluated there.
// Only conservativeness matters for correctness (a loose bound costs time, a wrong one would change the mosaic), so
// every step below errs outwards; tests/test_weights_first_host.py checks the function (this very code, through
// m2d_cell_weight_bounds) against real weight pyramids built by the oracle, nadir and tilted, both weight types.
//
// What is bounded (MultiBandMap2DCPU.cpp:396-425, 449-474): W_0(p) = wimg[rnd(phi(p))] for region px p whose rounded
// source position lies inside the frame, else 0, with wimg(s) = max(1e-5, 1 - |s - c| / dmax) (or its square) and phi
// the inverse homography; W_l = pyrDown^l(W_0): every W_l(u) is a convex combination (positive taps summing to 1;
// BORDER_REFLECT_101 folds taps back INTO the region) of W_0 at px within R_l = 2^(l+1) - 2 of q = u * 2^l.
//
// Bounds for cell (cx, cy) of level l (px u in [c*B, (c+1)*B) per axis, B = max(32 >> l, 1)):
//  * support rect S = kernel centres q(u) of the cell's px, grown by R_l, clipped to the region; phi(S) is the convex
//    quad of the 4 mapped corners (denominators positive), its bounding box grown by 1 px covers every rounded tap.
//  * LOOSE (always valid): W <= wimg at the box's point nearest to the frame centre (wimg decreases with the distance
//    to the centre; taps outside the frame contribute 0); W >= 0.  All of S outside the frame: W == 0.
//  * TIGHT (S maps inside the frame, so every tap is a real sample): x -> 1 - |phi(x) - c| / dmax is Lipschitz with
//    constant Lip / dmax, Lip >= |d phi| on S: phi - c = N_c / d with N_c affine and d linear, so
//    |d phi| <= (|N_c'|_2 + max_S |phi - c| * |d'|) / min_S d.  A pyrDown^l tap set has E|x - q| <= sqrt(2) sigma_l with
//    sigma_l^2 = (4^l - 1) / 3 (variance of the iterated [1 4 6 4 1]/16 kernel; folding taps at the region border only
//    moves them closer to q), and q ranges over the cell's centre rect (half diagonal hd).  Hence
//    | W_l(u) - w(phi(r)) | <= Lip / dmax * (hd + sqrt(2) sigma_l) + (0.7072 / dmax + float slack), r = rect centre,
//    where 0.7072 covers the nearest-neighbour rounding of the sample position.  Squaring (WeightType 1) and the 1e-5
//    clamp are monotone, so they are applied to the two ends.
//  * TIGHTER UPPER bound for WeightType 0 (same precondition): s -> 1 - |s - c| / dmax is CONCAVE, so by Jensen
//    sum_i k_i w(phi(q_i)) <= w(sum_i k_i phi(q_i)); the tap set is symmetric about q (no tap is folded: S lies inside
//    the frame, hence inside the region), so the mean of the mapped taps differs from phi(q) only by the curvature of
//    the projective map: |sum_i k_i phi(q_i) - phi(q)| <= 1/2 M2 E|x - q|^2 = M2 sigma_l^2 with
//    M2 = 2 sqrt(2) Lip |d'| / min d >= |d^2 phi| (phi d = N is affine, so d^2 phi = -(d phi (x) d' + d' (x) d phi) / d).
//    The sqrt(2) sigma_l term of the Lipschitz bound -- 26 px at level 5 -- shrinks to a fraction of a px.
#pragma once
#include <math.h>

#include "geom.h"

namespace m2d {

// region px -> source px with the FP32 copy of the inverse homography
M2D_HD void bounds_project(const float* m, float x, float y, float& sx, float& sy, float& den) {
    den = m[6] * x + m[7] * y + m[8];
    const float r = 1.f / den;
    sx = (m[0] * x + m[1] * y + m[2]) * r;
    sy = (m[3] * x + m[4] * y + m[5]) * r;
}

// m: FP32 inverse homography; nx, ny: frame region in tiles; sw, sh: source size; l: pyramid level; (cx, cy): cell in
// REGION coordinates (0 <= cx < nx*8).  On return lo <= W_l(u) <= hi for every px u of the cell.
M2D_HD void cell_weight_bounds(const float* m, int nx, int ny, int sw, int sh, int weight_type, int l, int cx, int cy,
                               float* lo_out, float* hi_out) {
    const float kSqrt2Sigma[6] = {0.f, 1.4143f, 3.1623f, 6.4808f, 13.0385f, 26.1152f};
    const int B = (32 >> l) > 0 ? (32 >> l) : 1;
    const int R = l == 0 ? 0 : (2 << l) - 2;
    const int rw = nx * 256, rh = ny * 256;
    const int qx0 = (cx * B) << l, qx1 = ((cx + 1) * B - 1) << l, qy0 = (cy * B) << l, qy1 = ((cy + 1) * B - 1) << l;
    const float sx0 = (float)(qx0 - R > 0 ? qx0 - R : 0), sx1 = (float)(qx1 + R < rw - 1 ? qx1 + R : rw - 1);
    const float sy0 = (float)(qy0 - R > 0 ? qy0 - R : 0), sy1 = (float)(qy1 + R < rh - 1 ? qy1 + R : rh - 1);
    float px[4], py[4], dn[4];
    bounds_project(m, sx0, sy0, px[0], py[0], dn[0]);
    bounds_project(m, sx1, sy0, px[1], py[1], dn[1]);
    bounds_project(m, sx0, sy1, px[2], py[2], dn[2]);
    bounds_project(m, sx1, sy1, px[3], py[3], dn[3]);
    const float dmin = fminf(fminf(dn[0], dn[1]), fminf(dn[2], dn[3]));
    if (!(dmin > 1e-3f)) { *lo_out = 0.f; *hi_out = INFINITY; return; }   // cannot reason about this frame here: never cull it
    const float bx0 = fminf(fminf(px[0], px[1]), fminf(px[2], px[3])) - 1.f, bx1 = fmaxf(fmaxf(px[0], px[1]), fmaxf(px[2], px[3])) + 1.f;
    const float by0 = fminf(fminf(py[0], py[1]), fminf(py[2], py[3])) - 1.f, by1 = fmaxf(fmaxf(py[0], py[1]), fmaxf(py[2], py[3])) + 1.f;
    const float fw = (float)sw, fh = (float)sh;
    if (bx1 < -0.5f || bx0 > fw - 0.5f || by1 < -0.5f || by0 > fh - 0.5f) { *lo_out = 0.f; *hi_out = 0.f; return; }   // samples outside the frame only
    const float xc = (float)(sw / 2), yc = (float)(sh / 2);
    const float dmax = sqrtf(xc * xc + yc * yc);
    // loose upper bound: the weight image at the box's point nearest to the frame centre
    const float ndx = fmaxf(0.f, fmaxf(bx0 - xc, xc - bx1)), ndy = fmaxf(0.f, fmaxf(by0 - yc, yc - by1));
    const float ndis = 1.f - fminf(sqrtf(ndx * ndx + ndy * ndy) / dmax, 1.f);
    const float nv = weight_type == 0 ? ndis : ndis * ndis;
    const float hi_loose = fmaxf(nv, 1e-5f) * (1.f + 3e-5f) + 1e-7f;
    const bool inside = bx0 >= 0.5f && bx1 <= fw - 1.5f && by0 >= 0.5f && by1 <= fh - 1.5f;
    if (!inside) { *lo_out = 0.f; *hi_out = hi_loose; return; }
    // Lipschitz constant of phi - c on S
    const float a = m[0] - xc * m[6], b = m[1] - xc * m[7], c = m[3] - yc * m[6], d = m[4] - yc * m[7];
    const float fro2 = a * a + b * b + c * c + d * d, det = a * d - b * c;
    const float disc = fmaxf(fro2 * fro2 - 4.f * det * det, 0.f);
    const float smax = sqrtf(0.5f * (fro2 + sqrtf(disc)));   // largest singular value of N_c'
    float rho2 = 0.f;
    for (int i = 0; i < 4; i++) rho2 = fmaxf(rho2, (px[i] - xc) * (px[i] - xc) + (py[i] - yc) * (py[i] - yc));
    const float lip = (smax + sqrtf(rho2) * sqrtf(m[6] * m[6] + m[7] * m[7])) / dmin * 1.001f;
    // the cell's kernel-centre rect: its centre r and half diagonal hd
    const float rx = 0.5f * (float)(qx0 + qx1), ry = 0.5f * (float)(qy0 + qy1);
    float csx, csy, cden;
    bounds_project(m, rx, ry, csx, csy, cden);
    const float rr = sqrtf((csx - xc) * (csx - xc) + (csy - yc) * (csy - yc));
    const float ddx = (float)(qx1 - qx0), ddy = (float)(qy1 - qy0);
    const float hd = 0.5f * sqrtf(ddx * ddx + ddy * ddy);
    const float spread = (lip * (hd + kSqrt2Sigma[l]) + 0.7072f) / dmax;
    const float base = 1.f - rr / dmax;
    float w_hi = fminf(base + spread, 1.f), w_lo = base - spread;
    if (weight_type != 0) {
        w_hi = fmaxf(w_hi, 0.f); w_hi = w_hi * w_hi;
        w_lo = fmaxf(w_lo, 0.f); w_lo = w_lo * w_lo;
    } else {
        // concavity: the kernel spread only enters through the curvature of the projective map
        const float sig2 = 0.5f * kSqrt2Sigma[l] * kSqrt2Sigma[l];
        const float m2 = 2.8285f * lip * sqrtf(m[6] * m[6] + m[7] * m[7]) / dmin;
        const float spread_j = (lip * hd + m2 * sig2 + 0.7072f) / dmax + 1e-5f;
        w_hi = fminf(w_hi, base + spread_j);
    }
    const float hi = fmaxf(w_hi, 1e-5f) * (1.f + 3e-5f) + 2e-6f;
    const float lo = fmaxf(fmaxf(w_lo, 1e-5f) * (1.f - 3e-5f) - 2e-6f, 0.f);
    *lo_out = lo;
    *hi_out = fminf(hi, hi_loose);
}

// ---------------------------------------------------------------------------------------------------------
// Pull mode (kernels_wf.cu mbs_mark / mbs_pull): which source px the image warp of one 32 x 32 level-0 cell can read.
// The cell's region px (X0..X0+31, Y0..Y0+31) map through the inverse homography to a convex quad (positive
// denominators); mb_sample4 reads, for every px, the taps floor(f) and floor(f)+1 per axis of a coordinate rounded to
// 1/32 px, folded back into the frame by BORDER_REFLECT (index -k reads k-1, index n-1+k reads n-k).  Returns the
// inclusive source rectangle [lox, hix] x [loy, hiy] that contains every such tap: the quad's bounding box, -2 / +3 px,
// with whatever sticks out of the frame reflected back in (the whole axis if it sticks out by more than the frame).
// false: the denominator is not positive somewhere on the cell (or a coordinate is absurd) -> the caller takes the
// whole frame.  Only conservativeness matters; tests/test_weights_first_host.py checks this very code (through
// m2d_pull_cell_rect) against a brute-force walk of the taps.
M2D_HD bool pull_cell_rect(const double* hinv, int X0, int Y0, int sw, int sh, int* lox, int* hix, int* loy, int* hiy) {
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    bool bad = false;
    for (int k = 0; k < 4; k++) {
        const double X = (double)(X0 + ((k & 1) ? 31 : 0)), Y = (double)(Y0 + ((k & 2) ? 31 : 0));
        const double W = hinv[6] * X + hinv[7] * Y + hinv[8];
        const double fx = (hinv[0] * X + hinv[1] * Y + hinv[2]) / W, fy = (hinv[3] * X + hinv[4] * Y + hinv[5]) / W;
        if (!(W > 0.0) || !(fabs(fx) < 1e8) || !(fabs(fy) < 1e8)) bad = true;
        x0 = fmin(x0, fx); x1 = fmax(x1, fx); y0 = fmin(y0, fy); y1 = fmax(y1, fy);
    }
    if (bad) { *lox = 0; *hix = sw - 1; *loy = 0; *hiy = sh - 1; return false; }
    const int sx0 = (int)floor(x0) - 2, sx1 = (int)floor(x1) + 3, sy0 = (int)floor(y0) - 2, sy1 = (int)floor(y1) + 3;
    int lx = sx0 > 0 ? sx0 : 0, hx = sx1 < sw - 1 ? sx1 : sw - 1, ly = sy0 > 0 ? sy0 : 0, hy = sy1 < sh - 1 ? sy1 : sh - 1;
    if (sx0 < 0) { lx = 0; const int r = -sx0 - 1 < sw - 1 ? -sx0 - 1 : sw - 1; hx = hx > r ? hx : r; }
    if (sx1 >= sw) { hx = sw - 1; int r = 2 * sw - 1 - sx1; r = r > 0 ? r : 0; lx = lx < r ? lx : r; }
    if (sy0 < 0) { ly = 0; const int r = -sy0 - 1 < sh - 1 ? -sy0 - 1 : sh - 1; hy = hy > r ? hy : r; }
    if (sy1 >= sh) { hy = sh - 1; int r = 2 * sh - 1 - sy1; r = r > 0 ? r : 0; ly = ly < r ? ly : r; }
    if (sx0 < -sw || sx1 >= 2 * sw || lx > hx) { lx = 0; hx = sw - 1; }
    if (sy0 < -sh || sy1 >= 2 * sh || ly > hy) { ly = 0; hy = sh - 1; }
    *lox = lx; *hix = hx; *loy = ly; *hiy = hy;
    return true;
}

}  // namespace m2d
