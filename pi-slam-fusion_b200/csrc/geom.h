// geom.h — FP64 geometry of the feed() path, shared verbatim by host code and the bounds kernel.
//
// Everything here must evaluate to the same bits on the CPU and on the GPU, because it decides WHICH tiles a
// frame touches and the 1/32-px sampling grid (tile footprint parity is a bit-exact requirement).  The file is
// therefore compiled with nvcc --fmad=false (device) and -ffp-contract=off (host): every a*b+c below is a
// rounded multiply followed by a rounded add, as in the reference's SSE2 build (reference CMakeLists.txt:17-30).
// IEEE division, sqrt, floor and ceil are correctly rounded on both sides.
//
// Restates: GSLAM/core/SO3.h:435-450,481-484, SE3.h:70-89 (pose algebra); Map2D.h:53-64 (UnProject);
// Map2DCPU.cpp:44-92 / MultiBandMap2DCPU.cpp:199-255 (grid layout); Map2DCPU.cpp:163-233,284-298 (bounds +
// homography); Map2DCPU.cpp:339-382 (spreadMap); cv::getPerspectiveTransform + cv::invert 3x3 (OpenCV).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define M2D_HD __host__ __device__ __forceinline__
#else
#define M2D_HD inline
#endif

namespace m2d {

struct Quat { double x, y, z, w; };
struct Vec3 { double x, y, z; };
struct Pose { Quat r; Vec3 t; };

M2D_HD Quat qmul(const Quat& l, const Quat& r) {
    Quat o;
    o.x = l.w * r.x + l.x * r.w + l.y * r.z - l.z * r.y;
    o.y = l.w * r.y + l.y * r.w + l.z * r.x - l.x * r.z;
    o.z = l.w * r.z + l.z * r.w + l.x * r.y - l.y * r.x;
    o.w = l.w * r.w - l.x * r.x - l.y * r.y - l.z * r.z;
    return o;
}
M2D_HD Quat qinv(const Quat& q) { Quat o; o.x = -q.x; o.y = -q.y; o.z = -q.z; o.w = q.w; return o; }
M2D_HD Vec3 qrot(const Quat& q, const Vec3& p) {
    Quat sp; sp.x = p.x; sp.y = p.y; sp.z = p.z; sp.w = 0;
    sp = qmul(qmul(q, sp), qinv(q));
    Vec3 o; o.x = sp.x; o.y = sp.y; o.z = sp.z;
    return o;
}
M2D_HD Pose pose_inverse(const Pose& p) {
    Pose o; o.r = qinv(p.r);
    Vec3 v = qrot(o.r, p.t);
    o.t.x = -v.x; o.t.y = -v.y; o.t.z = -v.z;
    return o;
}
M2D_HD Pose pose_mul(const Pose& a, const Pose& b) {
    Pose o; o.r = qmul(a.r, b.r);
    Vec3 rt = qrot(a.r, b.t);
    o.t.x = a.t.x + rt.x; o.t.y = a.t.y + rt.y; o.t.z = a.t.z + rt.z;
    return o;
}
M2D_HD Pose pose_from7(const double* v) {
    Pose p; p.t.x = v[0]; p.t.y = v[1]; p.t.z = v[2]; p.r.x = v[3]; p.r.y = v[4]; p.r.z = v[5]; p.r.w = v[6];
    return p;
}

// Everything the bounds computation needs; plain data so it can be passed to a kernel by value.
struct GridGeom {
    double cam_w, cam_h, cx, cy, fxinv, fyinv;  // Map2DPrepare::_camera, _fxinv, _fyinv
    Pose plane_inv;                              // plane.inverse(), applied in feed() (Map2DCPU.cpp:136)
    double min_x, min_y, max_x, max_y;           // Map2DCPUData::_min/_max (x,y)
    double ele_size, ele_size_inv, length_pixel_inv;
    int w, h;                                    // tiles
};

struct FrameBounds {
    int ok;                  // 0: rejected (oblique view / singular homography)
    int x0, y0, x1, y1;      // xminInt, yminInt, xmaxInt, ymaxInt against GridGeom (may lie outside [0,w]x[0,h])
    double gx0, gy0, gx1, gy1;  // un-snapped ground bbox (xmin,ymin,xmax,ymax): spreadMap needs it
    double hinv[9];          // inverse homography: region px -> source px
};

// cv::getPerspectiveTransform: 8x8 system from float products, LU with partial pivoting (OpenCV hal::LU64f).
M2D_HD bool perspective_from_points(const float* src, const float* dst, double* M) {
    double A[8][8], b[8];
    for (int i = 0; i < 4; ++i) {
        float sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
        for (int c = 0; c < 8; c++) { A[i][c] = 0; A[i + 4][c] = 0; }
        A[i][0] = A[i + 4][3] = sx;
        A[i][1] = A[i + 4][4] = sy;
        A[i][2] = A[i + 4][5] = 1;
        float p0 = -sx * dx, p1 = -sy * dx, p2 = -sx * dy, p3 = -sy * dy;  // float products, as in OpenCV
        A[i][6] = p0; A[i][7] = p1; A[i + 4][6] = p2; A[i + 4][7] = p3;
        b[i] = dx; b[i + 4] = dy;
    }
    const double eps = 2.220446049250313e-16 * 100;
    for (int i = 0; i < 8; i++) {
        int k = i;
        for (int j = i + 1; j < 8; j++)
            if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
        if (fabs(A[k][i]) < eps) return false;
        if (k != i) {
            for (int j = i; j < 8; j++) { double t = A[i][j]; A[i][j] = A[k][j]; A[k][j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = -1 / A[i][i];
        for (int j = i + 1; j < 8; j++) {
            double alpha = A[j][i] * d;
            for (int c = i + 1; c < 8; c++) A[j][c] += alpha * A[i][c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = 7; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < 8; k++) s -= A[i][k] * b[k];
        b[i] = s / A[i][i];
    }
    for (int i = 0; i < 8; i++) M[i] = b[i];
    M[8] = 1.0;
    return true;
}

// cv::invert, 3x3 double closed form (what cv::warpPerspective applies to the forward matrix).
M2D_HD bool invert3x3(const double* s, double* t) {
    double c00 = s[4] * s[8] - s[5] * s[7], c01 = s[3] * s[8] - s[5] * s[6], c02 = s[3] * s[7] - s[4] * s[6];
    double d = s[0] * c00 - s[1] * c01 + s[2] * c02;
    if (d == 0.) return false;
    d = 1. / d;
    t[0] = c00 * d;
    t[1] = (s[2] * s[7] - s[1] * s[8]) * d;
    t[2] = (s[1] * s[5] - s[2] * s[4]) * d;
    t[3] = (s[5] * s[6] - s[3] * s[8]) * d;
    t[4] = (s[0] * s[8] - s[2] * s[6]) * d;
    t[5] = (s[2] * s[3] - s[0] * s[5]) * d;
    t[6] = c02 * d;
    t[7] = (s[1] * s[6] - s[0] * s[7]) * d;
    t[8] = (s[0] * s[4] - s[1] * s[3]) * d;
    return true;
}

// renderFrame part 1: corner rays, obliqueness test, ground hits, bbox, tile range, homography.
// Tile indices are computed against `g` as it is; the caller applies spreadMap (host) when the bbox leaves it
// and then calls again.  with_homography = false stops after the tile range (ok, x0..y1, gx0..gy1 are final, hinv stays 0):
// all a pose-only feed or a dry run needs, at a third of the cost.
M2D_HD void frame_bounds(const GridGeom& g, const double* pose7, FrameBounds* out, bool with_homography = true) {
    out->ok = 0;
    out->x0 = out->y0 = out->x1 = out->y1 = -1;
    for (int i = 0; i < 9; i++) out->hinv[i] = 0;
    Pose f = pose_mul(g.plane_inv, pose_from7(pose7));
    double ipx[4], ipy[4], px[4], py[4];
    ipx[0] = 0; ipy[0] = 0; ipx[1] = g.cam_w; ipy[1] = 0; ipx[2] = 0; ipy[2] = g.cam_h; ipx[3] = g.cam_w; ipy[3] = g.cam_h;
    double down_z = (f.t.z < 0) ? 1.0 : -1.0;
    for (int i = 0; i < 4; i++) {
        Vec3 u; u.x = (ipx[i] - g.cx) * g.fxinv; u.y = (ipy[i] - g.cy) * g.fyinv; u.z = 1.;
        Vec3 axis = qrot(f.r, u);
        // axis.dot(downLook) with downLook = (0,0,+-1): x*0 + y*0 + z*dz
        if (axis.x * 0.0 + axis.y * 0.0 + axis.z * down_z < 0.4) return;
        double s = f.t.z / axis.z;
        px[i] = f.t.x - axis.x * s;
        py[i] = f.t.y - axis.y * s;
    }
    double xmin = px[0], xmax = xmin, ymin = py[0], ymax = ymin;
    for (int i = 1; i < 4; i++) {
        if (px[i] < xmin) xmin = px[i];
        if (py[i] < ymin) ymin = py[i];
        if (px[i] > xmax) xmax = px[i];
        if (py[i] > ymax) ymax = py[i];
    }
    out->gx0 = xmin; out->gy0 = ymin; out->gx1 = xmax; out->gy1 = ymax;
    int xi0 = (int)floor((xmin - g.min_x) * g.ele_size_inv), yi0 = (int)floor((ymin - g.min_y) * g.ele_size_inv);
    int xi1 = (int)ceil((xmax - g.min_x) * g.ele_size_inv), yi1 = (int)ceil((ymax - g.min_y) * g.ele_size_inv);
    if (!with_homography) {
        // The full path also rejects a frame whose homography cannot be solved (LU pivot below 100 eps), which only happens
        // when the four ground hits are (nearly) collinear, e.g. a camera centre ON the plane.  Every shard must take the
        // same decision, so a quad that is anywhere near degenerate goes through the full path after all.
        const double area2 = fabs((px[3] - px[0]) * (py[2] - py[1]) - (px[2] - px[1]) * (py[3] - py[0]));   // 2 * area, from the diagonals
        const double ext = (xmax - xmin) * (xmax - xmin) + (ymax - ymin) * (ymax - ymin);
        if (!(area2 > 1e-6 * ext)) with_homography = true;
    }
    if (with_homography) {
        double ox = g.min_x + g.ele_size * xi0, oy = g.min_y + g.ele_size * yi0;
        float srcp[8], dstp[8];
        for (int i = 0; i < 4; i++) {
            srcp[2 * i] = (float)ipx[i]; srcp[2 * i + 1] = (float)ipy[i];
            dstp[2 * i] = (float)((px[i] - ox) * g.length_pixel_inv);
            dstp[2 * i + 1] = (float)((py[i] - oy) * g.length_pixel_inv);
        }
        double M[9];
        if (!perspective_from_points(srcp, dstp, M)) return;
        if (!invert3x3(M, out->hinv)) return;
    }
    out->x0 = xi0; out->y0 = yi0; out->x1 = xi1; out->y1 = yi1;
    out->ok = 1;
}

}  // namespace m2d
