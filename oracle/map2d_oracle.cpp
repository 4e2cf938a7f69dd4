/*
 * map2d_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A dependency-free restatement of the reference's Map2DCPU / MultiBandMap2DCPU feed() path and of the
 * OpenCV primitives it delegates to.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this; the product (pi-slam-fusion_b200/) never does.
 *
 * PARITY PIN STATUS: the reference ships no Map2D test, golden image or fixture, and its
 * arithmetic lives in an un-vendored third-party dependency, OpenCV 2.4.9 (README.md:21-22,
 * thirdparty/opencv-2.4.9.zip is a missing blob).  The reference itself cannot be compiled here (needs
 * OpenCV C++/Qt4/GL headers).  => "parity unpinned" by the reference's own tests.  What this oracle IS pinned
 * to (tests/test_oracle_vs_cv2.py, tests/golden/): the real OpenCV primitives as shipped in cv2 4.13 —
 * bit-exact for every integer primitive (warp 8UC4 / 16SC3, pyrDown/pyrUp 16S, Laplace/restore), bit-exact
 * for the nearest f32 warp, and within 2 ulp for the f32 pyrDown (whose float association differs between
 * OpenCV 2.4.9 and 4.x; this file follows 2.4.9's, see pyr_down_f32).
 *
 * Every function cites the reference lines (relative to /root/reference) or the OpenCV 2.4.9 routine it
 * restates.  Build: oracle/Makefile (g++ -O3 -ffp-contract=off, no -march: the reference was plain x86-64
 * SSE2 code, CMakeLists.txt:17-30, so no FMA contraction anywhere).
 */
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "../include/map2d_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

int g_threads = 1; /* the reference runs one worker thread (Map2DCPU.cpp:397-413) */
int g_f32_mode = 0; /* 0: OpenCV 2.4.9 float association in pyrDown(f32) (the reference); 1: OpenCV 4.x with
                       128-bit universal intrinsics (what cv2 4.13 executes) — used ONLY to pin against cv2 */

// ---------------------------------------------------------------------------------------------------
// Pose algebra — GSLAM/GSLAM/core/SO3.h:435-450,481-484 and SE3.h:70-89 (double precision).
// ---------------------------------------------------------------------------------------------------
struct Quat { double x, y, z, w; };
struct Vec3 { double x, y, z; };
struct Pose { Quat r; Vec3 t; };

inline Quat qmul(const Quat& l, const Quat& rq) {  // SO3::operator*(SO3)  SO3.h:435-442
    return Quat{l.w * rq.x + l.x * rq.w + l.y * rq.z - l.z * rq.y,
                l.w * rq.y + l.y * rq.w + l.z * rq.x - l.x * rq.z,
                l.w * rq.z + l.z * rq.w + l.x * rq.y - l.y * rq.x,
                l.w * rq.w - l.x * rq.x - l.y * rq.y - l.z * rq.z};
}
inline Quat qinv(const Quat& q) { return Quat{-q.x, -q.y, -q.z, q.w}; }  // SO3.h:481-484
inline Vec3 qrot(const Quat& q, const Vec3& p) {                          // SO3.h:445-450
    Quat sp{p.x, p.y, p.z, 0};
    sp = qmul(qmul(q, sp), qinv(q));
    return Vec3{sp.x, sp.y, sp.z};
}
inline Pose pose_inverse(const Pose& p) {  // SE3.h:70-73
    Quat ri = qinv(p.r);
    Vec3 v = qrot(ri, p.t);
    return Pose{ri, Vec3{-v.x, -v.y, -v.z}};
}
inline Pose pose_mul(const Pose& a, const Pose& b) {  // SE3.h:84-89
    Vec3 rt = qrot(a.r, b.t);
    return Pose{qmul(a.r, b.r), Vec3{a.t.x + rt.x, a.t.y + rt.y, a.t.z + rt.z}};
}
inline Pose pose_from7(const double* v) { return Pose{Quat{v[3], v[4], v[5], v[6]}, Vec3{v[0], v[1], v[2]}}; }

// ---------------------------------------------------------------------------------------------------
// OpenCV 2.4.9 primitives (imgproc/imgwarp.cpp, imgproc/pyramids.cpp, core/lapack.cpp)
// ---------------------------------------------------------------------------------------------------
inline int cv_round(double v) { return (int)lrint(v); }  // cvRound: _mm_cvtsd_si32, round-half-even
inline short sat_short(int v) { return (short)std::min(std::max(v, (int)SHRT_MIN), (int)SHRT_MAX); }
inline uint8_t sat_u8(int v) { return (uint8_t)std::min(std::max(v, 0), 255); }

// cv::borderInterpolate, BORDER_REFLECT (delta=0) / BORDER_REFLECT_101 (delta=1)
inline int border_reflect(int p, int len, int delta) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1 + delta;
        else p = len - 1 - (p - len) - delta;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// cv::getPerspectiveTransform — imgwarp.cpp; 8x8 system in double built from FLOAT products, solved by
// Gaussian elimination with partial pivoting (cv::LU as used by OpenCV >= 3; 2.4.9 passes DECOMP_SVD, which
// differs in the last ulps only — SURVEY.md §8c).
bool get_perspective_transform(const float* src, const float* dst, double* M) {
    double A[8][8], b[8];
    for (int i = 0; i < 4; ++i) {
        float sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
        A[i][0] = A[i + 4][3] = sx;
        A[i][1] = A[i + 4][4] = sy;
        A[i][2] = A[i + 4][5] = 1;
        A[i][3] = A[i][4] = A[i][5] = A[i + 4][0] = A[i + 4][1] = A[i + 4][2] = 0;
        A[i][6] = (float)(-sx * dx);
        A[i][7] = (float)(-sy * dx);
        A[i + 4][6] = (float)(-sx * dy);
        A[i + 4][7] = (float)(-sy * dy);
        b[i] = dx;
        b[i + 4] = dy;
    }
    const int m = 8;
    const double eps = 2.220446049250313e-16 * 100;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (std::fabs(A[j][i]) > std::fabs(A[k][i])) k = j;
        if (std::fabs(A[k][i]) < eps) return false;
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A[i][j], A[k][j]);
            std::swap(b[i], b[k]);
        }
        double d = -1 / A[i][i];
        for (int j = i + 1; j < m; j++) {
            double alpha = A[j][i] * d;
            for (int c = i + 1; c < m; c++) A[j][c] += alpha * A[i][c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < m; k++) s -= A[i][k] * b[k];
        b[i] = s / A[i][i];
    }
    for (int i = 0; i < 8; i++) M[i] = b[i];
    M[8] = 1.0;
    return true;
}

// cv::invert for 3x3 double (core/lapack.cpp, closed-form cofactor branch)
bool invert3x3(const double* s, double* t) {
#define S(r, c) s[(r)*3 + (c)]
    double d = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) +
               S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
    if (d == 0.) return false;
    d = 1. / d;
    t[0] = (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * d;
    t[1] = (S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * d;
    t[2] = (S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * d;
    t[3] = (S(1, 2) * S(2, 0) - S(1, 0) * S(2, 2)) * d;
    t[4] = (S(0, 0) * S(2, 2) - S(0, 2) * S(2, 0)) * d;
    t[5] = (S(0, 2) * S(1, 0) - S(0, 0) * S(1, 2)) * d;
    t[6] = (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0)) * d;
    t[7] = (S(0, 1) * S(2, 0) - S(0, 0) * S(2, 1)) * d;
    t[8] = (S(0, 0) * S(1, 1) - S(0, 1) * S(1, 0)) * d;
#undef S
    return true;
}

// Coordinate generation of cv::warpPerspectiveInvoker (imgwarp.cpp): per (bw0 x bh0) block the row base
// X0 = M0*x + M1*y + M2 is formed at the block's first column x and advanced by M0*x1 inside the block.
// `M` is the INVERSE map (dst -> src).  scale = 32 (INTER_LINEAR, 5 fractional bits) or 1 (INTER_NEAREST).
// Outputs the un-shifted integer coordinates X,Y for one destination row.
void warp_row_coords(const double* M, int y, int width, int height, double scale, int* X, int* Y) {
    const int BLOCK_SZ = 32;
    int bh0 = std::min(BLOCK_SZ / 2, height);
    int bw0 = std::min(BLOCK_SZ * BLOCK_SZ / bh0, width);
    for (int x = 0; x < width; x += bw0) {
        int bw = std::min(bw0, width - x);
        double X0 = M[0] * x + M[1] * y + M[2];
        double Y0 = M[3] * x + M[4] * y + M[5];
        double W0 = M[6] * x + M[7] * y + M[8];
        for (int x1 = 0; x1 < bw; x1++) {
            double W = W0 + M[6] * x1;
            W = W ? scale / W : 0;
            double fX = std::max((double)INT_MIN, std::min((double)INT_MAX, (X0 + M[0] * x1) * W));
            double fY = std::max((double)INT_MIN, std::min((double)INT_MAX, (Y0 + M[3] * x1) * W));
            X[x + x1] = cv_round(fX);
            Y[x + x1] = cv_round(fY);
        }
    }
}

// warpPerspective(8UC4, INTER_LINEAR, BORDER_CONSTANT 0) — remapBilinear<FixedPtCast<int,uchar,15>,...,short>:
// weights BilinearTab_i = tab_f*32768 (exact here: tab_f has 10 fractional bits), (sum + 16384) >> 15.
void warp_u8c4_linear_const0(const uint8_t* src, int sh, int sw, size_t sstep, const double* Minv, uint8_t* dst,
                             int dh, int dw) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < dh; y++) {
        std::vector<int> X(dw), Y(dw);
        warp_row_coords(Minv, y, dw, dh, 32.0, X.data(), Y.data());
        uint8_t* D = dst + (size_t)y * dw * 4;
        for (int x = 0; x < dw; x++, D += 4) {
            int sx = sat_short(X[x] >> 5), sy = sat_short(Y[x] >> 5);
            int a = X[x] & 31, b = Y[x] & 31;
            int w00 = (32 - a) * (32 - b) * 32, w01 = a * (32 - b) * 32, w10 = (32 - a) * b * 32, w11 = a * b * 32;
            if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) {
                D[0] = D[1] = D[2] = D[3] = 0;
                continue;
            }
            bool x0in = (unsigned)sx < (unsigned)sw, x1in = (unsigned)(sx + 1) < (unsigned)sw;
            bool y0in = (unsigned)sy < (unsigned)sh, y1in = (unsigned)(sy + 1) < (unsigned)sh;
            const uint8_t* r0 = src + (size_t)sy * sstep;
            const uint8_t* r1 = src + (size_t)(sy + 1) * sstep;
            for (int c = 0; c < 4; c++) {
                int v00 = (y0in && x0in) ? r0[sx * 4 + c] : 0;
                int v01 = (y0in && x1in) ? r0[(sx + 1) * 4 + c] : 0;
                int v10 = (y1in && x0in) ? r1[sx * 4 + c] : 0;
                int v11 = (y1in && x1in) ? r1[(sx + 1) * 4 + c] : 0;
                D[c] = sat_u8((v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15);
            }
        }
    }
}

// warpPerspective(16SC3, INTER_LINEAR, BORDER_REFLECT) — remapBilinear<Cast<float,short>,RemapNoVec,float>:
// float weights BilinearTab_f[b][a] = {(1-b/32)*(1-a/32), (1-b/32)*(a/32), (b/32)*(1-a/32), (b/32)*(a/32)},
// sum evaluated left to right in float, saturate_cast<short>(float) = cvRound.
void warp_s16c3_linear_reflect(const int16_t* src, int sh, int sw, const double* Minv, int16_t* dst, int dh, int dw) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < dh; y++) {
        std::vector<int> X(dw), Y(dw);
        warp_row_coords(Minv, y, dw, dh, 32.0, X.data(), Y.data());
        int16_t* D = dst + (size_t)y * dw * 3;
        for (int x = 0; x < dw; x++, D += 3) {
            int sx = sat_short(X[x] >> 5), sy = sat_short(Y[x] >> 5);
            int a = X[x] & 31, b = Y[x] & 31;
            float fa = a * (1.f / 32), fb = b * (1.f / 32);
            float w0 = (1.f - fb) * (1.f - fa), w1 = (1.f - fb) * fa, w2 = fb * (1.f - fa), w3 = fb * fa;
            int sx0 = border_reflect(sx, sw, 0), sx1 = border_reflect(sx + 1, sw, 0);
            int sy0 = border_reflect(sy, sh, 0), sy1 = border_reflect(sy + 1, sh, 0);
            const int16_t* p00 = src + ((size_t)sy0 * sw + sx0) * 3;
            const int16_t* p01 = src + ((size_t)sy0 * sw + sx1) * 3;
            const int16_t* p10 = src + ((size_t)sy1 * sw + sx0) * 3;
            const int16_t* p11 = src + ((size_t)sy1 * sw + sx1) * 3;
            for (int c = 0; c < 3; c++) {
                float v = p00[c] * w0 + p01[c] * w1 + p10[c] * w2 + p11[c] * w3;
                D[c] = sat_short((int)lrintf(v));
            }
        }
    }
}

// warpPerspective(32FC1, INTER_NEAREST, BORDER_CONSTANT 0) — remapNearest<float>
void warp_f32_nearest_const0(const float* src, int sh, int sw, const double* Minv, float* dst, int dh, int dw) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < dh; y++) {
        std::vector<int> X(dw), Y(dw);
        warp_row_coords(Minv, y, dw, dh, 1.0, X.data(), Y.data());
        float* D = dst + (size_t)y * dw;
        for (int x = 0; x < dw; x++) {
            int sx = sat_short(X[x]), sy = sat_short(Y[x]);
            D[x] = ((unsigned)sx < (unsigned)sw && (unsigned)sy < (unsigned)sh) ? src[(size_t)sy * sw + sx] : 0.f;
        }
    }
}

// cv::pyrDown — pyramids.cpp pyrDown_<FixedPtCast<int,short,8>,...>: separable [1 4 6 4 1], BORDER_REFLECT_101,
// integer accumulate, (sum + 128) >> 8.  dst = ((rows+1)/2) x ((cols+1)/2).
void pyr_down_s16(const int16_t* src, int rows, int cols, int cn, int16_t* dst) {
    int orows = (rows + 1) / 2, ocols = (cols + 1) / 2;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < orows; y++) {
        std::vector<int> acc((size_t)ocols * cn, 0);
        static const int kv[5] = {1, 4, 6, 4, 1};
        for (int dy = -2; dy <= 2; dy++) {
            const int16_t* S = src + (size_t)border_reflect(2 * y + dy, rows, 1) * cols * cn;
            for (int x = 0; x < ocols; x++) {
                int xm2 = border_reflect(2 * x - 2, cols, 1) * cn, xm1 = border_reflect(2 * x - 1, cols, 1) * cn;
                int x0 = 2 * x * cn, xp1 = border_reflect(2 * x + 1, cols, 1) * cn, xp2 = border_reflect(2 * x + 2, cols, 1) * cn;
                for (int c = 0; c < cn; c++) {
                    int h = S[x0 + c] * 6 + (S[xm1 + c] + S[xp1 + c]) * 4 + S[xm2 + c] + S[xp2 + c];
                    acc[(size_t)x * cn + c] += kv[dy + 2] * h;
                }
            }
        }
        int16_t* D = dst + (size_t)y * ocols * cn;
        for (int i = 0; i < ocols * cn; i++) D[i] = sat_short((acc[i] + 128) >> 8);
    }
}

// cv::pyrDown for CV_32F, OpenCV 2.4.9 association (x86-64 always has SSE):
//   horizontal (scalar code):  row = s[2x]*6 + (s[2x-1] + s[2x+1])*4 + s[2x-2] + s[2x+2]      (left to right)
//   vertical (PyrDownVec_32f): ((r0 + r4) + (r2 + r2)) + ((r1 + r3) + r2)*4,  then * (1/256)
// (the SSE body takes 8 columns per iteration; the remaining ocols % 8 columns use the scalar expression
//  r2*6 + (r1 + r3)*4 + r0 + r4.  Tile-path levels of <= 5 bands are multiples of 8 wide; deeper levels and
//  Map2DRender's sub-images are not.)
void pyr_down_f32(const float* src, int rows, int cols, float* dst) {
    int orows = (rows + 1) / 2, ocols = (cols + 1) / 2;
    // OpenCV 4.x (g_f32_mode 1): columns [1, 1+4k) with 1+4k <= width0 go through PyrDownVecH (4 lanes):
    //   s[2x]*6 + ((s[2x-1]+s[2x+1])*4 + (s[2x-2]+s[2x+2])); the rest use the scalar expression.  Vertically the
    //   first (ocols/4)*4 columns use the vector form below, the tail the scalar form.
    int width0 = std::min((cols - 3) / 2 + 1, ocols);
    int hvec_end = (g_f32_mode == 1 && width0 > 1) ? 1 + ((width0 - 1) / 4) * 4 : 0;  // vector cols: [1, hvec_end)
    int vvec_end = (g_f32_mode == 1) ? (ocols / 4) * 4 : (ocols / 8) * 8;  // 2.4.9: PyrDownVec_32f takes 8 columns per step
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < orows; y++) {
        std::vector<float> h((size_t)5 * ocols);
        for (int dy = -2; dy <= 2; dy++) {
            const float* S = src + (size_t)border_reflect(2 * y + dy, rows, 1) * cols;
            float* H = h.data() + (size_t)(dy + 2) * ocols;
            for (int x = 0; x < ocols; x++) {
                float sm2 = S[border_reflect(2 * x - 2, cols, 1)], sm1 = S[border_reflect(2 * x - 1, cols, 1)];
                float s0 = S[2 * x], sp1 = S[border_reflect(2 * x + 1, cols, 1)], sp2 = S[border_reflect(2 * x + 2, cols, 1)];
                if (x >= 1 && x < hvec_end) H[x] = s0 * 6.f + ((sm1 + sp1) * 4.f + (sm2 + sp2));
                else H[x] = s0 * 6 + (sm1 + sp1) * 4 + sm2 + sp2;
            }
        }
        const float *r0 = h.data(), *r1 = r0 + ocols, *r2 = r1 + ocols, *r3 = r2 + ocols, *r4 = r3 + ocols;
        float* D = dst + (size_t)y * ocols;
        for (int x = 0; x < ocols; x++) {
            if (x < vvec_end) {
                float t0 = (r0[x] + r4[x]) + (r2[x] + r2[x]);
                float t1 = (r1[x] + r3[x]) + r2[x];
                D[x] = (t0 + t1 * 4.f) * (1.f / 256);
            } else {
                D[x] = (r2[x] * 6 + (r1[x] + r3[x]) * 4 + r0[x] + r4[x]) * (1.f / 256);
            }
        }
    }
}

// cv::pyrUp — pyramids.cpp pyrUp_<FixedPtCast<int,short,6>,...>, dst exactly 2x: per axis
// even: s[i-1] + 6 s[i] + s[i+1], odd: 4 (s[i] + s[i+1]); index -1 -> 1 (reflect-101), index n -> n-1 (replicate).
void pyr_up_s16(const int16_t* src, int rows, int cols, int cn, int16_t* dst) {
    int orows = rows * 2, ocols = cols * 2;
    auto lo = [](int i, int n) { return i < 0 ? (n > 1 ? 1 : 0) : i; };
    auto hi = [](int i, int n) { return i >= n ? n - 1 : i; };
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < rows; y++) {
        std::vector<int> hrow((size_t)3 * ocols * cn);
        for (int k = 0; k < 3; k++) {
            int sy = y - 1 + k;
            sy = sy < 0 ? lo(sy, rows) : hi(sy, rows);
            const int16_t* S = src + (size_t)sy * cols * cn;
            int* H = hrow.data() + (size_t)k * ocols * cn;
            for (int x = 0; x < cols; x++) {
                int xl = lo(x - 1, cols) * cn, xc = x * cn, xr = hi(x + 1, cols) * cn;
                for (int c = 0; c < cn; c++) {
                    H[(2 * x) * cn + c] = S[xl + c] + S[xc + c] * 6 + S[xr + c];
                    H[(2 * x + 1) * cn + c] = (S[xc + c] + S[xr + c]) * 4;
                }
            }
        }
        const int *r0 = hrow.data(), *r1 = r0 + (size_t)ocols * cn, *r2 = r1 + (size_t)ocols * cn;
        int16_t* D0 = dst + (size_t)(2 * y) * ocols * cn;
        int16_t* D1 = D0 + (size_t)ocols * cn;
        for (int i = 0; i < ocols * cn; i++) {
            D0[i] = sat_short((r0[i] + r1[i] * 6 + r2[i] + 32) >> 6);
            D1[i] = sat_short(((r1[i] + r2[i]) * 4 + 32) >> 6);
        }
    }
    (void)orows;
}

inline void sub_sat_s16(int16_t* a, const int16_t* b, size_t n) {  // cv::subtract CV_16S
    for (size_t i = 0; i < n; i++) a[i] = sat_short((int)a[i] - (int)b[i]);
}
inline void add_sat_s16(int16_t* a, const int16_t* b, size_t n) {  // cv::add CV_16S
    for (size_t i = 0; i < n; i++) a[i] = sat_short((int)a[i] + (int)b[i]);
}

struct Img16 { int rows = 0, cols = 0; std::vector<int16_t> d; };  // CV_16SC3
struct ImgF { int rows = 0, cols = 0; std::vector<float> d; };     // CV_32FC1

// cv::detail::createLaplacePyr, non-8U branch (stitching/blenders.cpp): pyr[0] aliases the image.
void create_laplace_pyr(Img16&& img, int levels, std::vector<Img16>& pyr) {
    pyr.resize(levels + 1);
    pyr[0] = std::move(img);
    for (int i = 0; i < levels; i++) {
        pyr[i + 1].rows = (pyr[i].rows + 1) / 2;
        pyr[i + 1].cols = (pyr[i].cols + 1) / 2;
        pyr[i + 1].d.resize((size_t)pyr[i + 1].rows * pyr[i + 1].cols * 3);
        pyr_down_s16(pyr[i].d.data(), pyr[i].rows, pyr[i].cols, 3, pyr[i + 1].d.data());
    }
    std::vector<int16_t> tmp;
    for (int i = 0; i < levels; i++) {
        tmp.resize(pyr[i].d.size());
        pyr_up_s16(pyr[i + 1].d.data(), pyr[i + 1].rows, pyr[i + 1].cols, 3, tmp.data());
        sub_sat_s16(pyr[i].d.data(), tmp.data(), tmp.size());
    }
}
// cv::detail::restoreImageFromLaplacePyr
void restore_from_laplace_pyr(std::vector<Img16>& pyr) {
    std::vector<int16_t> tmp;
    for (size_t i = pyr.size() - 1; i > 0; --i) {
        tmp.resize(pyr[i - 1].d.size());
        pyr_up_s16(pyr[i].d.data(), pyr[i].rows, pyr[i].cols, 3, tmp.data());
        add_sat_s16(pyr[i - 1].d.data(), tmp.data(), tmp.size());
    }
}

// Weight images.  Map2DCPU.cpp:236-258 (8-bit alpha) and MultiBandMap2DCPU.cpp:396-418 (float).
// `sqrt` on floats under `using namespace std` is the float overload; w/2, h/2 are integer divisions.
void weight_image_u8(int w, int h, int weight_type, uint8_t* out) {
    float x_center = w / 2, y_center = h / 2;
    float dis_max = std::sqrt(x_center * x_center + y_center * y_center);
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            float dis = (i - y_center) * (i - y_center) + (j - x_center) * (j - x_center);
            dis = 1 - std::sqrt(dis) / dis_max;
            uint8_t a;
            if (0 == weight_type) a = (uint8_t)(dis * 254.);
            else a = (uint8_t)(dis * dis * 254);
            if (a < 2) a = 2;
            out[(size_t)i * w + j] = a;
        }
}
void weight_image_f32(int w, int h, int weight_type, float* out) {
    float x_center = w / 2, y_center = h / 2;
    float dis_max = std::sqrt(x_center * x_center + y_center * y_center);
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            float dis = (i - y_center) * (i - y_center) + (j - x_center) * (j - x_center);
            dis = 1 - std::sqrt(dis) / dis_max;
            float v = (0 == weight_type) ? dis : dis * dis;
            if (v <= 1e-5) v = 1e-5;  // double literal compare/assign, as written in the reference
            out[(size_t)i * w + j] = v;
        }
}

// ---------------------------------------------------------------------------------------------------
// The Map2D object (both modes)
// ---------------------------------------------------------------------------------------------------
struct Tile {
    std::vector<uint8_t> bgra;             // weighted: 256*256*4 (Map2DCPUEle::img)
    std::vector<std::vector<int16_t>> lap; // multi-band: pyr_laplace[level] (CV_16SC3)
    std::vector<std::vector<float>> wgt;   //             weights[level]     (CV_32FC1)
};

struct Map {
    int type = 0;
    m2d_config cfg{};
    int band_num = 5;
    bool valid = false;
    // Map2DPrepare
    double cam_w = 0, cam_h = 0, fx = 0, fy = 0, cx = 0, cy = 0, fxinv = 0, fyinv = 0;
    Pose plane{};
    // Map2DCPUData
    double ele_size = 0, ele_size_inv = 0, length_pixel = 0, length_pixel_inv = 0;
    Vec3 vmin{}, vmax{};
    int w = 0, h = 0;
    std::vector<std::shared_ptr<Tile>> data;
    // weight image cache
    int wimg_w = 0, wimg_h = 0;
    std::vector<uint8_t> wimg_u8;
    std::vector<float> wimg_f32;
    int last_rect[4] = {-1, -1, -1, -1};
    m2d_stats stats{};
    int org_x = 0, org_y = 0;  // absolute tile coordinate of slot (0,0): test stand-in for the sharded CUDA library
    bool pose_only_violation = false;
    int shard_origin = 0;

    bool owns(int tx, int ty) const {  // same rule as include/map2d_b200.h m2d_config.shard_*
        if (cfg.shard_count <= 1) return true;
        int a = ((cfg.shard_axis == 0) ? tx + org_x : ty + org_y) - shard_origin;
        int span = cfg.shard_span > 0 ? cfg.shard_span : 1;
        int q = (a >= 0) ? a / span : -((-a + span - 1) / span);
        int r = q % cfg.shard_count;
        if (r < 0) r += cfg.shard_count;
        return r == cfg.shard_rank;
    }
    size_t tile_bytes() const {
        if (type != M2D_TYPE_MULTIBAND) return (size_t)M2D_ELE_PIXELS * M2D_ELE_PIXELS * 4;
        size_t b = 0;
        for (int l = 0; l <= band_num; l++) { size_t n = (size_t)(M2D_ELE_PIXELS >> l); b += n * n * (3 * sizeof(int16_t) + sizeof(float)); }
        return b;
    }

    Vec3 unproject(double u, double v) const { return Vec3{(u - cx) * fxinv, (v - cy) * fyinv, 1.}; }  // Map2D.h:60-64

    bool prepare(const double* plane7, const double* cam6, int n, const double* poses);
    bool spread(double xmin, double ymin, double xmax, double ymax);
    bool feed(const uint8_t* bgr, int fw, int fh, size_t stride, const double* pose7);
    bool render_weighted(const uint8_t* bgr, size_t stride, const double* Minv, int x0, int y0, int x1, int y1);
    bool render_multiband(const uint8_t* bgr, size_t stride, const double* Minv, int x0, int y0, int x1, int y1);
    bool tile_bbox(int& minx, int& miny, int& maxx, int& maxy) const;
};

// Map2DPrepare::prepare (Map2D.cpp:32-49) + Map2DCPUData::prepare (Map2DCPU.cpp:44-92) /
// MultiBandMap2DCPUData::prepare (MultiBandMap2DCPU.cpp:199-255)
bool Map::prepare(const double* plane7, const double* cam6, int n, const double* poses) {
    if (n <= 0 || cam6[0] <= 0 || cam6[1] <= 0 || cam6[2] == 0 || cam6[3] == 0) return false;
    cam_w = cam6[0]; cam_h = cam6[1]; fx = cam6[2]; fy = cam6[3]; cx = cam6[4]; cy = cam6[5];
    fxinv = 1. / fx; fyinv = 1. / fy;
    plane = pose_from7(plane7);
    Vec3 mx{-1e10, -1e10, -1e10}, mn{1e10, 1e10, 1e10};
    Pose pinv = pose_inverse(plane);
    for (int i = 0; i < n; i++) {
        Pose p = pose_mul(pinv, pose_from7(poses + 7 * i));
        const Vec3& t = p.t;
        mx.x = t.x > mx.x ? t.x : mx.x; mx.y = t.y > mx.y ? t.y : mx.y; mx.z = t.z > mx.z ? t.z : mx.z;
        mn.x = t.x < mn.x ? t.x : mn.x; mn.y = t.y < mn.y ? t.y : mn.y; mn.z = t.z < mn.z ? t.z : mn.z;
    }
    if (mn.z * mx.z <= 0) return false;
    double hgt;
    if (type == M2D_TYPE_MULTIBAND) hgt = (mx.z > 0) ? mx.z : -mn.z;  // maxh, MultiBandMap2DCPU.cpp:222-224
    else hgt = (mn.z > 0) ? mn.z : -mx.z;                               // minh, Map2DCPU.cpp:67-69
    Vec3 a = unproject(cam_w, cam_h), b = unproject(0, 0);
    Vec3 line{a.x - b.x, a.y - b.y, a.z - b.z};
    double radius = 0.5 * hgt * std::sqrt((line.x * line.x + line.y * line.y));
    double lp = 0;
    if (type == M2D_TYPE_MULTIBAND) lp = cfg.resolution;
    if (!lp) {
        lp = 2 * radius / std::sqrt(cam_w * cam_w + cam_h * cam_h);
        lp /= cfg.scale;
    }
    length_pixel = lp;
    length_pixel_inv = 1. / lp;
    mn = Vec3{mn.x - radius, mn.y - radius, mn.z - 0};
    mx = Vec3{mx.x + radius, mx.y + radius, mx.z + 0};
    Vec3 center{0.5 * (mn.x + mx.x), 0.5 * (mn.y + mx.y), 0.5 * (mn.z + mx.z)};
    mn = Vec3{2 * mn.x - center.x, 2 * mn.y - center.y, 2 * mn.z - center.z};
    mx = Vec3{2 * mx.x - center.x, 2 * mx.y - center.y, 2 * mx.z - center.z};
    ele_size = M2D_ELE_PIXELS * lp;
    ele_size_inv = 1. / ele_size;
    w = (int)std::ceil((mx.x - mn.x) / ele_size);
    h = (int)std::ceil((mx.y - mn.y) / ele_size);
    mx.x = mn.x + ele_size * w;
    mx.y = mn.y + ele_size * h;
    vmin = mn; vmax = mx;
    org_x = org_y = 0;
    data.assign((size_t)w * h, nullptr);
    wimg_w = wimg_h = 0;
    valid = true;
    return true;
}

// spreadMap — Map2DCPU.cpp:339-382 / MultiBandMap2DCPU.cpp:561-604
bool Map::spread(double xmin, double ymin, double xmax, double ymax) {
    int xminInt = (int)std::floor((xmin - vmin.x) * ele_size_inv);
    int yminInt = (int)std::floor((ymin - vmin.y) * ele_size_inv);
    int xmaxInt = (int)std::ceil((xmax - vmin.x) * ele_size_inv);
    int ymaxInt = (int)std::ceil((ymax - vmin.y) * ele_size_inv);
    xminInt = std::min(xminInt, 0); yminInt = std::min(yminInt, 0);
    xmaxInt = std::max(xmaxInt, w); ymaxInt = std::max(ymaxInt, h);
    int nw = xmaxInt - xminInt, nh = ymaxInt - yminInt;
    double nminx = vmin.x + ele_size * xminInt, nminy = vmin.y + ele_size * yminInt;
    double nmaxx = nminx + nw * ele_size, nmaxy = nminy + nh * ele_size;
    std::vector<std::shared_ptr<Tile>> copy((size_t)nw * nh);
    for (int x = 0; x < w; x++)
        for (int y = 0; y < h; y++) copy[(size_t)(x - xminInt) + (size_t)(y - yminInt) * nw] = data[(size_t)y * w + x];
    data.swap(copy);
    vmin.x = nminx; vmin.y = nminy; vmax.x = nmaxx; vmax.y = nmaxy;
    w = nw; h = nh;
    org_x += xminInt; org_y += yminInt;
    return true;
}

// feed() + renderFrame() part 1 (bounds) — Map2DCPU.cpp:127-233 ≡ MultiBandMap2DCPU.cpp:288-394,
// homography Map2DCPU.cpp:284-298.
bool Map::feed(const uint8_t* bgr, int fw, int fh, size_t stride, const double* pose7) {
    stats.frames_fed++;
    if (!valid) return false;
    Pose f = pose_mul(pose_inverse(plane), pose_from7(pose7));
    if (fw != cam_w || fh != cam_h) return false;
    const double ipx[4] = {0, cam_w, 0, cam_w}, ipy[4] = {0, 0, cam_h, cam_h};
    double px[4], py[4];
    Vec3 down{0, 0, -1};
    if (f.t.z < 0) down = Vec3{0, 0, 1};
    for (int i = 0; i < 4; i++) {
        Vec3 axis = qrot(f.r, unproject(ipx[i], ipy[i]));
        if (axis.x * down.x + axis.y * down.y + axis.z * down.z < 0.4) return false;
        double s = f.t.z / axis.z;
        axis = Vec3{f.t.x - axis.x * s, f.t.y - axis.y * s, f.t.z - axis.z * s};
        px[i] = axis.x; py[i] = axis.y;
    }
    double xmin = px[0], xmax = xmin, ymin = py[0], ymax = ymin;
    for (int i = 1; i < 4; i++) {
        if (px[i] < xmin) xmin = px[i];
        if (py[i] < ymin) ymin = py[i];
        if (px[i] > xmax) xmax = px[i];
        if (py[i] > ymax) ymax = py[i];
    }
    if (xmin < vmin.x || xmax > vmax.x || ymin < vmin.y || ymax > vmax.y)
        if (!spread(xmin, ymin, xmax, ymax)) return false;
    int xminInt = (int)std::floor((xmin - vmin.x) * ele_size_inv);
    int yminInt = (int)std::floor((ymin - vmin.y) * ele_size_inv);
    int xmaxInt = (int)std::ceil((xmax - vmin.x) * ele_size_inv);
    int ymaxInt = (int)std::ceil((ymax - vmin.y) * ele_size_inv);
    if (xminInt < 0 || yminInt < 0 || xmaxInt > w || ymaxInt > h || xminInt >= xmaxInt || yminInt >= ymaxInt) return false;
    xmin = vmin.x + ele_size * xminInt;
    ymin = vmin.y + ele_size * yminInt;
    float srcp[8], dstp[8];
    for (int i = 0; i < 4; i++) {
        srcp[2 * i] = (float)ipx[i]; srcp[2 * i + 1] = (float)ipy[i];
        dstp[2 * i] = (float)((px[i] - xmin) * length_pixel_inv);
        dstp[2 * i + 1] = (float)((py[i] - ymin) * length_pixel_inv);
    }
    double M[9], Mi[9];
    if (!get_perspective_transform(srcp, dstp, M)) return false;
    if (!invert3x3(M, Mi)) return false;  // warpPerspective inverts the forward map itself
    last_rect[0] = xminInt; last_rect[1] = yminInt; last_rect[2] = xmaxInt; last_rect[3] = ymaxInt;
    if (!bgr) {  // pose-only feed of a sharded run (m2d_feed_poses): grid decisions only
        for (int y = yminInt; y < ymaxInt; y++)
            for (int x = xminInt; x < xmaxInt; x++)
                if (owns(x, y)) pose_only_violation = true;
        return true;
    }
    stats.frames_fused++;
    stats.input_px += (uint64_t)fw * fh;
    if (type == M2D_TYPE_MULTIBAND) return render_multiband(bgr, stride, Mi, xminInt, yminInt, xmaxInt, ymaxInt);
    return render_weighted(bgr, stride, Mi, xminInt, yminInt, xmaxInt, ymaxInt);
}

// Map2DCPU::renderFrame part 2 — Map2DCPU.cpp:234-335
bool Map::render_weighted(const uint8_t* bgr, size_t stride, const double* Minv, int x0, int y0, int x1, int y1) {
    int fw = (int)cam_w, fh = (int)cam_h;
    if (wimg_w != fw || wimg_h != fh || wimg_u8.empty()) {
        wimg_u8.resize((size_t)fw * fh);
        weight_image_u8(fw, fh, cfg.weight_type, wimg_u8.data());
        wimg_w = fw; wimg_h = fh;
    }
    std::vector<uint8_t> src((size_t)fw * fh * 4);
    for (int i = 0; i < fh; i++) {
        const uint8_t* s = bgr + (size_t)i * stride;
        uint8_t* d = src.data() + (size_t)i * fw * 4;
        const uint8_t* a = wimg_u8.data() + (size_t)i * fw;
        for (int j = 0; j < fw; j++) { d[4 * j] = s[3 * j]; d[4 * j + 1] = s[3 * j + 1]; d[4 * j + 2] = s[3 * j + 2]; d[4 * j + 3] = a[j]; }
    }
    int dw = (x1 - x0) * M2D_ELE_PIXELS, dh = (y1 - y0) * M2D_ELE_PIXELS;
    std::vector<uint8_t> dst((size_t)dw * dh * 4);
    warp_u8c4_linear_const0(src.data(), fh, fw, (size_t)fw * 4, Minv, dst.data(), dh, dw);
    for (int x = x0; x < x1; x++)
        for (int y = y0; y < y1; y++) {
            if (!owns(x, y)) continue;
            std::shared_ptr<Tile>& ele = data[(size_t)y * w + x];
            bool fresh = false;
            if (!ele) ele = std::make_shared<Tile>();
            if (ele->bgra.empty()) { ele->bgra.assign((size_t)M2D_ELE_PIXELS * M2D_ELE_PIXELS * 4, 0); fresh = true; }
            uint8_t* eleP = ele->bgra.data();
            uint64_t wins = 0, foot = 0;
            for (int ey = 0; ey < M2D_ELE_PIXELS; ey++) {
                const uint8_t* dstP = dst.data() + ((size_t)((y - y0) * M2D_ELE_PIXELS + ey) * dw + (size_t)(x - x0) * M2D_ELE_PIXELS) * 4;
                for (int ex = 0; ex < M2D_ELE_PIXELS; ex++, dstP += 4, eleP += 4) {
                    foot += dstP[3] > 0;
                    if (eleP[3] < dstP[3]) { memcpy(eleP, dstP, 4); wins++; }
                }
            }
            stats.region_px[0] += (uint64_t)M2D_ELE_PIXELS * M2D_ELE_PIXELS;
            if (fresh) stats.fresh_px[0] += (uint64_t)M2D_ELE_PIXELS * M2D_ELE_PIXELS;
            if (!fresh) stats.win_px[0] += wins;  // wins on first-touch tiles are accounted as fresh_px
            stats.footprint_px += foot;
        }
    return true;
}

// MultiBandMap2DCPU::renderFrame part 2 — MultiBandMap2DCPU.cpp:395-557 (default CV_16SC3 path)
bool Map::render_multiband(const uint8_t* bgr, size_t stride, const double* Minv, int x0, int y0, int x1, int y1) {
    int fw = (int)cam_w, fh = (int)cam_h;
    if (wimg_w != fw || wimg_h != fh || wimg_f32.empty()) {
        wimg_f32.resize((size_t)fw * fh);
        weight_image_f32(fw, fh, cfg.weight_type, wimg_f32.data());
        wimg_w = fw; wimg_h = fh;
    }
    std::vector<int16_t> img_src((size_t)fw * fh * 3);  // frame.convertTo(CV_16SC3)
    for (int i = 0; i < fh; i++) {
        const uint8_t* s = bgr + (size_t)i * stride;
        int16_t* d = img_src.data() + (size_t)i * fw * 3;
        for (int j = 0; j < fw * 3; j++) d[j] = s[j];
    }
    int dw = (x1 - x0) * M2D_ELE_PIXELS, dh = (y1 - y0) * M2D_ELE_PIXELS;
    Img16 image_warped; image_warped.rows = dh; image_warped.cols = dw; image_warped.d.resize((size_t)dw * dh * 3);
    std::vector<ImgF> pyr_w(band_num + 1);
    pyr_w[0].rows = dh; pyr_w[0].cols = dw; pyr_w[0].d.resize((size_t)dw * dh);
    warp_s16c3_linear_reflect(img_src.data(), fh, fw, Minv, image_warped.d.data(), dh, dw);
    warp_f32_nearest_const0(wimg_f32.data(), fh, fw, Minv, pyr_w[0].d.data(), dh, dw);
    std::vector<Img16> pyr_l;
    create_laplace_pyr(std::move(image_warped), band_num, pyr_l);
    for (int i = 0; i < band_num; i++) {
        pyr_w[i + 1].rows = (pyr_w[i].rows + 1) / 2; pyr_w[i + 1].cols = (pyr_w[i].cols + 1) / 2;
        pyr_w[i + 1].d.resize((size_t)pyr_w[i + 1].rows * pyr_w[i + 1].cols);
        pyr_down_f32(pyr_w[i].d.data(), pyr_w[i].rows, pyr_w[i].cols, pyr_w[i + 1].d.data());
    }
    for (int x = x0; x < x1; x++)
        for (int y = y0; y < y1; y++) {
            if (!owns(x, y)) continue;
            std::shared_ptr<Tile>& ele = data[(size_t)y * w + x];
            if (!ele) ele = std::make_shared<Tile>();
            if (ele->lap.empty()) { ele->lap.resize(band_num + 1); ele->wgt.resize(band_num + 1); }
            int width = M2D_ELE_PIXELS, height = M2D_ELE_PIXELS;
            for (int i = 0; i <= band_num; ++i) {
                int pc = pyr_l[i].cols;
                size_t org = (size_t)(x - x0) * width + (size_t)(y - y0) * height * pc;
                stats.region_px[i] += (uint64_t)width * height;
                if (ele->lap[i].empty()) {  // fresh: copy the rect
                    ele->lap[i].resize((size_t)width * height * 3);
                    ele->wgt[i].resize((size_t)width * height);
                    for (int ey = 0; ey < height; ey++) {
                        memcpy(&ele->lap[i][(size_t)ey * width * 3], &pyr_l[i].d[(org + (size_t)ey * pc) * 3], (size_t)width * 3 * sizeof(int16_t));
                        memcpy(&ele->wgt[i][(size_t)ey * width], &pyr_w[i].d[org + (size_t)ey * pc], (size_t)width * sizeof(float));
                    }
                    stats.fresh_px[i] += (uint64_t)width * height;
                } else {
                    uint64_t wins = 0;
                    for (int ey = 0; ey < height; ey++) {
                        const int16_t* srcL = &pyr_l[i].d[(org + (size_t)ey * pc) * 3];
                        const float* srcW = &pyr_w[i].d[org + (size_t)ey * pc];
                        int16_t* dstL = &ele->lap[i][(size_t)ey * width * 3];
                        float* dstW = &ele->wgt[i][(size_t)ey * width];
                        for (int ex = 0; ex < width; ex++)
                            if (srcW[ex] >= dstW[ex]) {
                                dstL[3 * ex] = srcL[3 * ex]; dstL[3 * ex + 1] = srcL[3 * ex + 1]; dstL[3 * ex + 2] = srcL[3 * ex + 2];
                                dstW[ex] = srcW[ex];
                                wins++;
                            }
                    }
                    stats.win_px[i] += wins;
                }
                width /= 2; height /= 2;
            }
        }
    return true;
}

bool Map::tile_bbox(int& minx, int& miny, int& maxx, int& maxy) const {
    minx = miny = (int)1e6; maxx = maxy = (int)-1e6;
    for (int x = 0; x < w; x++)
        for (int y = 0; y < h; y++) {
            const auto& e = data[(size_t)x + (size_t)y * w];
            if (!e) continue;
            if (type == M2D_TYPE_MULTIBAND ? e->lap.empty() : e->bgra.empty()) continue;
            minx = std::min(minx, x); miny = std::min(miny, y);
            maxx = std::max(maxx, x); maxy = std::max(maxy, y);
        }
    if (maxx < minx) return false;
    maxx += 1; maxy += 1;
    return true;
}

}  // namespace

#define RENDER_ORACLE_PART 1
#include "render_oracle.inl"  // Map2DRender (type 4): warps, blender, renderFrames
#undef RENDER_ORACLE_PART

// ---------------------------------------------------------------------------------------------------
// extern "C" surface (mirrors include/map2d_b200.h with the orc_ prefix, plus the primitives)
// ---------------------------------------------------------------------------------------------------
extern "C" {

void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int orc_get_threads() { return g_threads; }
void orc_set_f32_mode(int m) { g_f32_mode = m; }

int orc_get_perspective_transform(const float* src8, const float* dst8, double* M9) { return get_perspective_transform(src8, dst8, M9) ? 0 : 1; }
int orc_invert3x3(const double* M, double* Mi) { return invert3x3(M, Mi) ? 0 : 1; }
// the three warps take the FORWARD matrix like cv::warpPerspective and invert it themselves
int orc_warp_u8c4(const uint8_t* src, int sh, int sw, const double* M, uint8_t* dst, int dh, int dw) {
    double Mi[9]; if (!invert3x3(M, Mi)) return 1;
    warp_u8c4_linear_const0(src, sh, sw, (size_t)sw * 4, Mi, dst, dh, dw); return 0;
}
int orc_warp_s16c3_reflect(const int16_t* src, int sh, int sw, const double* M, int16_t* dst, int dh, int dw) {
    double Mi[9]; if (!invert3x3(M, Mi)) return 1;
    warp_s16c3_linear_reflect(src, sh, sw, Mi, dst, dh, dw); return 0;
}
int orc_warp_f32_nearest(const float* src, int sh, int sw, const double* M, float* dst, int dh, int dw) {
    double Mi[9]; if (!invert3x3(M, Mi)) return 1;
    warp_f32_nearest_const0(src, sh, sw, Mi, dst, dh, dw); return 0;
}
void orc_pyrdown_s16(const int16_t* src, int rows, int cols, int cn, int16_t* dst) { pyr_down_s16(src, rows, cols, cn, dst); }
void orc_pyrdown_f32(const float* src, int rows, int cols, float* dst) { pyr_down_f32(src, rows, cols, dst); }
void orc_pyrup_s16(const int16_t* src, int rows, int cols, int cn, int16_t* dst) { pyr_up_s16(src, rows, cols, cn, dst); }
void orc_weight_image_u8(int w, int h, int weight_type, uint8_t* out) { weight_image_u8(w, h, weight_type, out); }
void orc_weight_image_f32(int w, int h, int weight_type, float* out) { weight_image_f32(w, h, weight_type, out); }

struct orc_map { Map m; RenderState rs; };

int orc_create(int type, const m2d_config* cfg, orc_map** out) {
    *out = nullptr;
    if (type == M2D_TYPE_GPU) type = M2D_TYPE_CPU;  // Map2D.cpp:57-65
    if (type != M2D_TYPE_CPU && type != M2D_TYPE_MULTIBAND && type != M2D_TYPE_RENDER) return M2D_ERR_UNSUPPORTED;
    if (cfg && cfg->force_float) return M2D_ERR_UNSUPPORTED;
    orc_map* o = new orc_map();
    o->m.type = type;
    if (cfg) o->m.cfg = *cfg;
    else { memset(&o->m.cfg, 0, sizeof(m2d_config)); o->m.cfg.scale = 1; o->m.cfg.band_number = 5; }
    if (o->m.cfg.scale == 0) o->m.cfg.scale = 1;
    int bn = o->m.cfg.band_number > 0 ? o->m.cfg.band_number : 5;
    o->m.band_num = std::min(bn, (int)std::ceil(std::log((double)M2D_ELE_PIXELS) / std::log(2.0)));  // MultiBandMap2DCPU.cpp:263
    *out = o;
    return M2D_OK;
}
void orc_destroy(orc_map* o) { delete o; }
int orc_prepare(orc_map* o, const double* plane, const double* cam, int n, const double* poses) {
    Map fresh; fresh.type = o->m.type; fresh.cfg = o->m.cfg; fresh.band_num = o->m.band_num;
    if (!fresh.prepare(plane, cam, n, poses)) return M2D_REJECTED;
    o->m = std::move(fresh);
    return M2D_OK;
}
int orc_feed(orc_map* o, const uint8_t* bgr, int w, int h, size_t stride, const double* pose) {
    if (o->m.type == M2D_TYPE_RENDER) return M2D_REJECTED;  // Map2DRender::renderFrame is `return false` (Map2DRender.cpp:464-467)
    return o->m.feed(bgr, w, h, stride, pose) ? M2D_OK : M2D_REJECTED;
}
int orc_plan_rects(orc_map* o, int n, const double* poses, int* rects) {  // stand-in for m2d_plan_rects
    if (!o->m.valid) return M2D_ERR_STATE;
    Map tmp = o->m;  // tiles are shared pointers: the copy is metadata only, and a pose-only feed never touches them
    for (int i = 0; i < n; i++) {
        int* r = rects + 4 * (size_t)i;
        tmp.last_rect[0] = tmp.last_rect[1] = tmp.last_rect[2] = tmp.last_rect[3] = -1;
        bool ok = tmp.feed(nullptr, (int)tmp.cam_w, (int)tmp.cam_h, 0, poses + 7 * (size_t)i);
        if (ok && tmp.last_rect[2] > tmp.last_rect[0]) {
            r[0] = tmp.last_rect[0] + tmp.org_x; r[1] = tmp.last_rect[1] + tmp.org_y;
            r[2] = tmp.last_rect[2] + tmp.org_x; r[3] = tmp.last_rect[3] + tmp.org_y;
        } else r[0] = r[1] = r[2] = r[3] = -1;
    }
    return M2D_OK;
}
int orc_set_shard(orc_map* o, int rank, int count, int axis, int span, int origin) {  // stand-in for m2d_set_shard
    if (count < 1 || rank < 0 || rank >= count || (axis != 0 && axis != 1) || span < 1) return M2D_ERR_ARG;
    for (const auto& e : o->m.data) if (e) return M2D_ERR_STATE;
    o->m.cfg.shard_rank = rank; o->m.cfg.shard_count = count; o->m.cfg.shard_axis = axis; o->m.cfg.shard_span = span;
    o->m.shard_origin = origin;
    return M2D_OK;
}
int orc_feed_poses(orc_map* o, int n, const double* poses, int* result) {  // stand-in for m2d_feed_poses
    int violations = 0;
    for (int i = 0; i < n; i++) {
        o->m.pose_only_violation = false;
        bool ok = o->m.feed(nullptr, (int)o->m.cam_w, (int)o->m.cam_h, 0, poses + 7 * (size_t)i);
        int st = !ok ? M2D_REJECTED : o->m.pose_only_violation ? M2D_ERR_ARG : M2D_OK;
        violations += st == M2D_ERR_ARG;
        if (result) result[i] = st;
    }
    return violations ? M2D_ERR_ARG : M2D_OK;
}
int orc_get_grid(orc_map* o, int* w, int* h, double* mn, double* mx, double* lp) {
    if (!o->m.valid) return M2D_ERR_STATE;
    *w = o->m.w; *h = o->m.h;
    mn[0] = o->m.vmin.x; mn[1] = o->m.vmin.y; mn[2] = o->m.vmin.z;
    mx[0] = o->m.vmax.x; mx[1] = o->m.vmax.y; mx[2] = o->m.vmax.z;
    *lp = o->m.length_pixel;
    return M2D_OK;
}
int orc_last_rect(orc_map* o, int* rect) { memcpy(rect, o->m.last_rect, sizeof(int) * 4); return M2D_OK; }
int orc_get_tile(orc_map* o, int tx, int ty, int level, void* out, float* weight) {
    Map& m = o->m;
    if (!m.valid || tx < 0 || ty < 0 || tx >= m.w || ty >= m.h) return M2D_ERR_ARG;
    const auto& e = m.data[(size_t)ty * m.w + tx];
    if (!e) return M2D_REJECTED;
    if (m.type == M2D_TYPE_MULTIBAND) {
        if (level < 0 || level > m.band_num || e->lap.empty()) return e->lap.empty() ? M2D_REJECTED : M2D_ERR_ARG;
        memcpy(out, e->lap[level].data(), e->lap[level].size() * sizeof(int16_t));
        if (weight) memcpy(weight, e->wgt[level].data(), e->wgt[level].size() * sizeof(float));
    } else {
        if (level != 0) return M2D_ERR_ARG;
        if (e->bgra.empty()) return M2D_REJECTED;
        memcpy(out, e->bgra.data(), e->bgra.size());
    }
    return M2D_OK;
}
// Map2DCPU::save (Map2DCPU.cpp:523-560) / MultiBandMap2DCPU::save (MultiBandMap2DCPU.cpp:779-841), in memory, over an explicit
// window of tiles (grid coordinates; absent tiles are the zeros the reference pastes), of which `crop` is written to `out`.
static int collapse_window(Map& m, uint8_t* out, const int win[4], const int crop[4], int* w, int* h, int* channels) {
    const int x0 = win[0], y0 = win[1], x1 = win[2], y1 = win[3];
    if (x1 <= x0 || y1 <= y0 || crop[0] < x0 || crop[1] < y0 || crop[2] > x1 || crop[3] > y1 || crop[2] <= crop[0] || crop[3] <= crop[1]) return M2D_ERR_ARG;
    int tw = x1 - x0, th = y1 - y0;
    int cn = m.type == M2D_TYPE_MULTIBAND ? 3 : 4;
    const size_t CW = (size_t)(crop[2] - crop[0]) * M2D_ELE_PIXELS, CH = (size_t)(crop[3] - crop[1]) * M2D_ELE_PIXELS;
    *w = (int)CW; *h = (int)CH; *channels = cn;
    if (!out) return M2D_OK;
    size_t W = (size_t)tw * M2D_ELE_PIXELS, H = (size_t)th * M2D_ELE_PIXELS;
    auto tile_at = [&](int x, int y) -> const Tile* {
        if (x < 0 || y < 0 || x >= m.w || y >= m.h) return nullptr;
        return m.data[(size_t)x + (size_t)y * m.w].get();
    };
    std::vector<uint8_t> full(W * H * cn);
    if (m.type != M2D_TYPE_MULTIBAND) {
        memset(full.data(), 0, full.size());  // the reference leaves untouched tiles uninitialised; we define 0
        for (int x = x0; x < x1; x++)
            for (int y = y0; y < y1; y++) {
                const Tile* e = tile_at(x, y);
                if (!e || e->bgra.empty()) continue;
                for (int ey = 0; ey < M2D_ELE_PIXELS; ey++)
                    memcpy(full.data() + (((size_t)(y - y0) * M2D_ELE_PIXELS + ey) * W + (size_t)(x - x0) * M2D_ELE_PIXELS) * 4,
                           &e->bgra[(size_t)ey * M2D_ELE_PIXELS * 4], (size_t)M2D_ELE_PIXELS * 4);
            }
    } else {
        int L = m.band_num;
        std::vector<Img16> pyr(L + 1);
        std::vector<float> w0(W * H, 0.f);
        for (int i = 0; i <= L; i++) {
            int n = M2D_ELE_PIXELS >> i;
            pyr[i].rows = th * n; pyr[i].cols = tw * n;
            pyr[i].d.assign((size_t)pyr[i].rows * pyr[i].cols * 3, 0);
        }
        for (int x = x0; x < x1; x++)
            for (int y = y0; y < y1; y++) {
                const Tile* e = tile_at(x, y);
                if (!e || e->lap.empty()) continue;
                for (int i = 0; i <= L; i++) {
                    int n = M2D_ELE_PIXELS >> i;
                    for (int ey = 0; ey < n; ey++) {
                        memcpy(&pyr[i].d[(((size_t)(y - y0) * n + ey) * pyr[i].cols + (size_t)(x - x0) * n) * 3], &e->lap[i][(size_t)ey * n * 3], (size_t)n * 3 * sizeof(int16_t));
                        if (i == 0) memcpy(&w0[((size_t)(y - y0) * n + ey) * W + (size_t)(x - x0) * n], &e->wgt[0][(size_t)ey * n], (size_t)n * sizeof(float));
                    }
                }
            }
        restore_from_laplace_pyr(pyr);
        uint8_t bg = sat_u8(m.cfg.background);
        size_t npx = W * H;
        for (size_t p = 0; p < npx; p++) {
            if (w0[p] == 0) { full[3 * p] = full[3 * p + 1] = full[3 * p + 2] = bg; continue; }
            full[3 * p] = sat_u8(pyr[0].d[3 * p]); full[3 * p + 1] = sat_u8(pyr[0].d[3 * p + 1]); full[3 * p + 2] = sat_u8(pyr[0].d[3 * p + 2]);
        }
    }
    for (size_t r = 0; r < CH; r++)
        memcpy(out + r * CW * cn, full.data() + (((size_t)(crop[1] - y0) * M2D_ELE_PIXELS + r) * W + (size_t)(crop[0] - x0) * M2D_ELE_PIXELS) * cn, CW * cn);
    return M2D_OK;
}
int orc_get_image(orc_map* o, uint8_t* out, int* w, int* h, int* channels, int* tile_min_x, int* tile_min_y) {
    Map& m = o->m;
    if (!m.valid || m.w == 0 || m.h == 0) return M2D_REJECTED;
    int x0, y0, x1, y1;
    if (!m.tile_bbox(x0, y0, x1, y1)) return M2D_REJECTED;
    *tile_min_x = x0; *tile_min_y = y0;
    const int win[4] = {x0, y0, x1, y1};
    return collapse_window(m, out, win, win, w, h, channels);
}
// stand-ins for m2d_tile_bbox / m2d_get_image_rect / m2d_export_tiles_rect / m2d_drop_tiles_rect (sharded save, host-logic tests)
int orc_tile_bbox(orc_map* o, int* bbox_abs) {
    Map& m = o->m;
    int x0, y0, x1, y1;
    if (!m.valid || !m.tile_bbox(x0, y0, x1, y1)) return M2D_REJECTED;
    bbox_abs[0] = x0 + m.org_x; bbox_abs[1] = y0 + m.org_y; bbox_abs[2] = x1 + m.org_x; bbox_abs[3] = y1 + m.org_y;
    return M2D_OK;
}
int orc_get_image_rect(orc_map* o, uint8_t* out, int, const int* window_abs, const int* crop_abs, int* w, int* h, int* channels) {
    Map& m = o->m;
    if (!m.valid) return M2D_ERR_STATE;
    const int win[4] = {window_abs[0] - m.org_x, window_abs[1] - m.org_y, window_abs[2] - m.org_x, window_abs[3] - m.org_y};
    const int crop[4] = {crop_abs[0] - m.org_x, crop_abs[1] - m.org_y, crop_abs[2] - m.org_x, crop_abs[3] - m.org_y};
    return collapse_window(m, out, win, crop, w, h, channels);
}
int orc_drop_tiles_rect(orc_map* o, const int* rect_abs, int* n_dropped) {
    Map& m = o->m;
    int n = 0;
    for (int y = std::max(rect_abs[1] - m.org_y, 0); y < std::min(rect_abs[3] - m.org_y, m.h); y++)
        for (int x = std::max(rect_abs[0] - m.org_x, 0); x < std::min(rect_abs[2] - m.org_x, m.w); x++) {
            auto& e = m.data[(size_t)y * m.w + x];
            if (e) { e.reset(); n++; }
        }
    if (n_dropped) *n_dropped = n;
    return M2D_OK;
}
size_t orc_tile_bytes(orc_map* o) { return o->m.tile_bytes(); }
int orc_tile_count(orc_map* o) {
    int n = 0;
    for (auto& e : o->m.data) if (e && !(o->m.type == M2D_TYPE_MULTIBAND ? e->lap.empty() : e->bgra.empty())) n++;
    return n;
}
int orc_export_tiles_rect(orc_map* o, const int* rect_abs, int max_tiles, int* abs_xy, uint8_t* dst, int, int* n_out);
int orc_export_tiles(orc_map* o, int max_tiles, int* abs_xy, uint8_t* dst, int, int* n_out) {
    return orc_export_tiles_rect(o, nullptr, max_tiles, abs_xy, dst, 0, n_out);
}
int orc_export_tiles_rect(orc_map* o, const int* rect_abs, int max_tiles, int* abs_xy, uint8_t* dst, int, int* n_out) {
    Map& m = o->m;
    int n = 0;
    size_t tb = m.tile_bytes();
    int ry0 = 0, ry1 = m.h, rx0 = 0, rx1 = m.w;
    if (rect_abs) {
        rx0 = std::max(rect_abs[0] - m.org_x, 0); ry0 = std::max(rect_abs[1] - m.org_y, 0);
        rx1 = std::min(rect_abs[2] - m.org_x, m.w); ry1 = std::min(rect_abs[3] - m.org_y, m.h);
    }
    for (int y = ry0; y < ry1; y++)
        for (int x = rx0; x < rx1; x++) {
            const auto& e = m.data[(size_t)y * m.w + x];
            if (!e || (m.type == M2D_TYPE_MULTIBAND ? e->lap.empty() : e->bgra.empty())) continue;
            if (max_tiles == 0) { n++; continue; }   // count only
            if (n >= max_tiles) return M2D_ERR_ARG;
            abs_xy[2 * n] = x + m.org_x; abs_xy[2 * n + 1] = y + m.org_y;
            uint8_t* d = dst + (size_t)n * tb;
            if (m.type != M2D_TYPE_MULTIBAND) memcpy(d, e->bgra.data(), tb);
            else for (int l = 0; l <= m.band_num; l++) {
                memcpy(d, e->lap[l].data(), e->lap[l].size() * sizeof(int16_t)); d += e->lap[l].size() * sizeof(int16_t);
                memcpy(d, e->wgt[l].data(), e->wgt[l].size() * sizeof(float)); d += e->wgt[l].size() * sizeof(float);
            }
            n++;
        }
    *n_out = n;
    return M2D_OK;
}
int orc_import_tiles(orc_map* o, int n, const int* abs_xy, const uint8_t* src, int) {
    Map& m = o->m;
    size_t tb = m.tile_bytes();
    for (int i = 0; i < n; i++) {
        int x = abs_xy[2 * i] - m.org_x, y = abs_xy[2 * i + 1] - m.org_y;
        if (x < 0 || y < 0 || x >= m.w || y >= m.h) return M2D_ERR_ARG;
        auto& e = m.data[(size_t)y * m.w + x];
        if (!e) e = std::make_shared<Tile>();
        const uint8_t* d = src + (size_t)i * tb;
        if (m.type != M2D_TYPE_MULTIBAND) e->bgra.assign(d, d + tb);
        else {
            e->lap.resize(m.band_num + 1); e->wgt.resize(m.band_num + 1);
            for (int l = 0; l <= m.band_num; l++) {
                size_t px = (size_t)(M2D_ELE_PIXELS >> l) * (M2D_ELE_PIXELS >> l);
                e->lap[l].assign((const int16_t*)d, (const int16_t*)d + px * 3); d += px * 3 * sizeof(int16_t);
                e->wgt[l].assign((const float*)d, (const float*)d + px); d += px * sizeof(float);
            }
        }
    }
    return M2D_OK;
}
// Display-time per-tile collapse — MultiBandMap2DCPUEle::blend + updateTexture (MultiBandMap2DCPU.cpp:77-188) as driven
// by draw() (:702-742): with HighQualityShow and all 9 neighbours present, every level is extended by a border of
// 1 << (levels-1-i) px taken from the neighbours, restored, and the centre 256x256 is cropped; otherwise the tile is
// restored alone.  Pixels with weights[0]==0 are zeroed, then 16S -> 8U.  Weighted mode shows the BGRA tile itself
// (Map2DCPU.cpp:497-505).
int orc_get_tile_image(orc_map* o, int tx, int ty, int high_quality, uint8_t* out, int* channels) {
    Map& m = o->m;
    if (!m.valid || tx < 0 || ty < 0 || tx >= m.w || ty >= m.h) return M2D_ERR_ARG;
    const auto& e = m.data[(size_t)ty * m.w + tx];
    if (!e) return M2D_REJECTED;
    const int E = M2D_ELE_PIXELS;
    if (m.type != M2D_TYPE_MULTIBAND) {
        if (e->bgra.empty()) return M2D_REJECTED;
        *channels = 4;
        memcpy(out, e->bgra.data(), e->bgra.size());
        return M2D_OK;
    }
    if (e->lap.empty()) return M2D_REJECTED;
    *channels = 3;
    const int L = m.band_num, levels = L + 1;
    bool all9 = high_quality != 0;
    const Tile* nb[9];
    for (int yi = ty - 1, k = 0; yi <= ty + 1; yi++)
        for (int xi = tx - 1; xi <= tx + 1; xi++, k++) {
            const Tile* t = (yi < 0 || yi >= m.h || xi < 0 || xi >= m.w) ? nullptr : m.data[(size_t)yi * m.w + xi].get();
            if (!t || t->lap.empty()) all9 = false;
            nb[k] = t;
        }
    std::vector<Img16> pyr(levels);
    int b0 = 0;
    for (int i = 0; i < levels; i++) {
        int n = E >> i, b = all9 ? (1 << (levels - i - 1)) : 0, d = n + 2 * b;
        if (i == 0) b0 = b;
        pyr[i].rows = pyr[i].cols = d;
        pyr[i].d.assign((size_t)d * d * 3, 0);
        for (int y = 0; y < 3; y++)
            for (int x = 0; x < 3; x++) {
                if (!all9 && !(x == 1 && y == 1)) continue;
                const Tile* t = all9 ? nb[3 * y + x] : e.get();
                int sw_ = (x == 1) ? n : b, sh_ = (y == 1) ? n : b;
                int sx = (x == 0) ? (n - b) : 0, sy = (y == 0) ? (n - b) : 0;
                int dx = (x == 0) ? 0 : ((x == 1) ? b : (d - b)), dy = (y == 0) ? 0 : ((y == 1) ? b : (d - b));
                for (int r = 0; r < sh_; r++)
                    memcpy(&pyr[i].d[((size_t)(dy + r) * d + dx) * 3], &t->lap[i][((size_t)(sy + r) * n + sx) * 3], (size_t)sw_ * 3 * sizeof(int16_t));
            }
    }
    restore_from_laplace_pyr(pyr);
    const int d0 = pyr[0].cols;
    for (int y = 0; y < E; y++)
        for (int x = 0; x < E; x++) {
            const int16_t* s = &pyr[0].d[((size_t)(y + b0) * d0 + (x + b0)) * 3];
            uint8_t* q = out + ((size_t)y * E + x) * 3;
            if (e->wgt[0][(size_t)y * E + x] == 0) { q[0] = q[1] = q[2] = 0; continue; }
            q[0] = sat_u8(s[0]); q[1] = sat_u8(s[1]); q[2] = sat_u8(s[2]);
        }
    return M2D_OK;
}
int orc_get_stats(orc_map* o, m2d_stats* out) { *out = o->m.stats; return M2D_OK; }

// Bounds only, against the current grid, without spreadMap (what m2d_compute_bounds returns).
int orc_compute_bounds(orc_map* o, int n, const double* poses, int* rects, double* hinv) {
    Map& m = o->m;
    if (!m.valid) return M2D_ERR_STATE;
    for (int k = 0; k < n; k++) {
        int* r = rects + 4 * k; double* Hi = hinv + 9 * k;
        r[0] = r[1] = r[2] = r[3] = -1;
        for (int i = 0; i < 9; i++) Hi[i] = 0;
        Pose f = pose_mul(pose_inverse(m.plane), pose_from7(poses + 7 * k));
        const double ipx[4] = {0, m.cam_w, 0, m.cam_w}, ipy[4] = {0, 0, m.cam_h, m.cam_h};
        double px[4], py[4];
        Vec3 down{0, 0, -1};
        if (f.t.z < 0) down = Vec3{0, 0, 1};
        bool ok = true;
        for (int i = 0; i < 4 && ok; i++) {
            Vec3 axis = qrot(f.r, m.unproject(ipx[i], ipy[i]));
            if (axis.x * down.x + axis.y * down.y + axis.z * down.z < 0.4) { ok = false; break; }
            double s = f.t.z / axis.z;
            px[i] = f.t.x - axis.x * s; py[i] = f.t.y - axis.y * s;
        }
        if (!ok) continue;
        double xmin = px[0], xmax = xmin, ymin = py[0], ymax = ymin;
        for (int i = 1; i < 4; i++) {
            if (px[i] < xmin) xmin = px[i];
            if (py[i] < ymin) ymin = py[i];
            if (px[i] > xmax) xmax = px[i];
            if (py[i] > ymax) ymax = py[i];
        }
        int xi0 = (int)std::floor((xmin - m.vmin.x) * m.ele_size_inv), yi0 = (int)std::floor((ymin - m.vmin.y) * m.ele_size_inv);
        int xi1 = (int)std::ceil((xmax - m.vmin.x) * m.ele_size_inv), yi1 = (int)std::ceil((ymax - m.vmin.y) * m.ele_size_inv);
        double ox = m.vmin.x + m.ele_size * xi0, oy = m.vmin.y + m.ele_size * yi0;
        float srcp[8], dstp[8];
        for (int i = 0; i < 4; i++) {
            srcp[2 * i] = (float)ipx[i]; srcp[2 * i + 1] = (float)ipy[i];
            dstp[2 * i] = (float)((px[i] - ox) * m.length_pixel_inv);
            dstp[2 * i + 1] = (float)((py[i] - oy) * m.length_pixel_inv);
        }
        double M[9];
        if (!get_perspective_transform(srcp, dstp, M) || !invert3x3(M, Hi)) continue;
        r[0] = xi0; r[1] = yi0; r[2] = xi1; r[3] = yi1;
    }
    return M2D_OK;
}

#define RENDER_ORACLE_PART 2
#include "render_oracle.inl"  // orc_render_*: stand-ins for m2d_render_*
#undef RENDER_ORACLE_PART

}  // extern "C"
