/*
 * render_oracle.inl — CPU ORACLE of Map2DRender (Map2D type 4).  TEST INFRASTRUCTURE ONLY.
 * Included at the end of map2d_oracle.cpp (same translation unit: it uses that file's OpenCV primitives).
 *
 * Restates Map2DFusion/Map2DRender.cpp: the batch driver renderFrames (:479-760) and its inline copy of OpenCV's
 * MultiBandBlender (:52-310) -- without the GUI (cv::imshow / cv::waitKey, Map2DRender.ShowPyrLaplace) and without the
 * seam finder (Map2DRender.EnableSeam = 0; cv::detail::DpSeamFinder is a sequential dynamic program outside the
 * data-parallel path).  Three blends:
 *   0  what the reference executes: CV_32F weights, per level `if (w >= dst_w) { dst_w = w; dst = src; }`   (:206-213)
 *   1  the `#else` branch = stock OpenCV, CV_32F weights: dst += short(src * w), dst_w += w, normalise      (:214-220,264-267)
 *   2  the CV_16S branch  = stock OpenCV, CV_16S weights: dst += short((src * w) >> 8), dst_w += w, normalise (:227-249)
 * Blends 1 and 2 are pinned bit-exactly against the REAL cv2.detail_MultiBandBlender (tests/test_render_oracle.py), blend 0
 * against the same loop rebuilt from cv2 primitives.
 */
#if RENDER_ORACLE_PART == 1
namespace {

// warpPerspective(8UC3, INTER_LINEAR, BORDER_REFLECT) -- Map2DRender.cpp:586.  Same fixed-point bilinear as the 8UC4 warp
// above (remapBilinear<FixedPtCast<int,uchar,15>, RemapVec_8u, short>), taps fetched through cv::borderInterpolate.
void warp_u8c3_linear_reflect(const uint8_t* src, int sh, int sw, size_t sstep, const double* Minv, uint8_t* dst, int dh, int dw) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < dh; y++) {
        std::vector<int> X(dw), Y(dw);
        warp_row_coords(Minv, y, dw, dh, 32.0, X.data(), Y.data());
        uint8_t* D = dst + (size_t)y * dw * 3;
        for (int x = 0; x < dw; x++, D += 3) {
            int sx = sat_short(X[x] >> 5), sy = sat_short(Y[x] >> 5);
            int a = X[x] & 31, b = Y[x] & 31;
            int w00 = (32 - a) * (32 - b) * 32, w01 = a * (32 - b) * 32, w10 = (32 - a) * b * 32, w11 = a * b * 32;
            int sx0 = border_reflect(sx, sw, 0), sx1 = border_reflect(sx + 1, sw, 0);
            int sy0 = border_reflect(sy, sh, 0), sy1 = border_reflect(sy + 1, sh, 0);
            const uint8_t *r0 = src + (size_t)sy0 * sstep, *r1 = src + (size_t)sy1 * sstep;
            for (int c = 0; c < 3; c++)
                D[c] = sat_u8((r0[sx0 * 3 + c] * w00 + r0[sx1 * 3 + c] * w01 + r1[sx0 * 3 + c] * w10 + r1[sx1 * 3 + c] * w11 + (1 << 14)) >> 15);
        }
    }
}

// warpPerspective(8UC1, INTER_NEAREST, BORDER_CONSTANT 0) -- Map2DRender.cpp:587 (remapNearest<uchar>)
void warp_u8c1_nearest_const0(const uint8_t* src, int sh, int sw, const double* Minv, uint8_t* dst, int dh, int dw) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < dh; y++) {
        std::vector<int> X(dw), Y(dw);
        warp_row_coords(Minv, y, dw, dh, 1.0, X.data(), Y.data());
        uint8_t* D = dst + (size_t)y * dw;
        for (int x = 0; x < dw; x++) {
            int sx = sat_short(X[x]), sy = sat_short(Y[x]);
            D[x] = ((unsigned)sx < (unsigned)sw && (unsigned)sy < (unsigned)sh) ? src[(size_t)sy * sw + sx] : 0;
        }
    }
}

// The weight image of Map2DRender::renderFrames, Map2DRender.cpp:507-529: centre at w*0.5 (not w/2), squared falloff,
// float arithmetic, `*p = dis*dis*254` truncated to a byte, floor 1.
void render_weight_image_u8(int w, int h, uint8_t* out) {
    float x_center = w * 0.5, y_center = h * 0.5;
    float dis_maxInv = 1. / std::sqrt(x_center * x_center + y_center * y_center);
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            float dis = (i - y_center) * (i - y_center) + (j - x_center) * (j - x_center);
            dis = 1 - std::sqrt(dis) * dis_maxInv;
            uint8_t p = (uint8_t)(dis * dis * 254);
            if (p < 1) p = 1;
            out[(size_t)i * w + j] = p;
        }
}

struct RenderFrame {
    bool ok = false;
    int iw = 0, ih = 0;          // sizes[idx]
    float cwx = 0, cwy = 0;      // cornersWorld[idx] (cv::Point2f)
    int cx = 0, cy = 0;          // cornersImages[idx]
    std::vector<uint8_t> img;    // imgwarped[idx]  (8UC3)
    std::vector<uint8_t> mask;   // maskwarped[idx] (8UC1)
};

struct RenderState {
    bool have = false;
    int W = 0, H = 0;            // dst_roi_final_ (the blended canvas: whole tiles)
    int num_bands = 0;
    int tile_x0 = 0, tile_y0 = 0;  // absolute tile coordinate of the canvas origin
    std::vector<int16_t> result;   // CV_16SC3, masked px zeroed (Blender::blend)
    std::vector<uint8_t> mask;     // dst_mask: 255 where the level-0 weight exceeds WEIGHT_EPS
    std::vector<RenderFrame> frames;
};

// Map2DFusion::MultiBandBlender (Map2DRender.cpp:52-310; == cv::detail::MultiBandBlender of OpenCV 2.4.9 except for the
// `#if 1` selection in feed()).
struct RenderBlender {
    int blend = 0, actual_num_bands = 5, num_bands = 0;
    int W = 0, H = 0, Wf = 0, Hf = 0;          // dst_roi_ (padded), dst_roi_final_
    std::vector<Img16> dst_lap;                // dst_pyr_laplace_
    std::vector<ImgF> dst_wf;                  // dst_band_weights_ (CV_32F)
    std::vector<std::vector<int16_t>> dst_ws;  // dst_band_weights_ (CV_16S)

    void prepare(int w, int h) {  // :73-101
        Wf = w; Hf = h;
        double max_len = (double)std::max(w, h);
        num_bands = std::min(actual_num_bands, (int)std::ceil(std::log(max_len) / std::log(2.0)));
        w += ((1 << num_bands) - w % (1 << num_bands)) % (1 << num_bands);
        h += ((1 << num_bands) - h % (1 << num_bands)) % (1 << num_bands);
        W = w; H = h;
        dst_lap.assign(num_bands + 1, Img16());
        dst_wf.assign(num_bands + 1, ImgF());
        dst_ws.assign(num_bands + 1, std::vector<int16_t>());
        int r = h, c = w;
        for (int i = 0; i <= num_bands; i++) {
            if (i) { r = (r + 1) / 2; c = (c + 1) / 2; }
            dst_lap[i].rows = r; dst_lap[i].cols = c; dst_lap[i].d.assign((size_t)r * c * 3, 0);
            dst_wf[i].rows = r; dst_wf[i].cols = c;
            if (blend == 2) dst_ws[i].assign((size_t)r * c, 0);
            else dst_wf[i].d.assign((size_t)r * c, 0.f);
        }
    }

    void feed(const uint8_t* img, const uint8_t* mask, int iw, int ih, int tlx, int tly) {  // :102-255
        int gap = 3 * (1 << num_bands);
        int tnx = std::max(0, tlx - gap), tny = std::max(0, tly - gap);
        int bnx = std::min(W, tlx + iw + gap), bny = std::min(H, tly + ih + gap);
        tnx = (tnx >> num_bands) << num_bands;
        tny = (tny >> num_bands) << num_bands;
        int width = bnx - tnx, height = bny - tny;
        width += ((1 << num_bands) - width % (1 << num_bands)) % (1 << num_bands);
        height += ((1 << num_bands) - height % (1 << num_bands)) % (1 << num_bands);
        bnx = tnx + width; bny = tny + height;
        int dy = std::max(bny - H, 0), dx = std::max(bnx - W, 0);
        tnx -= dx; bnx -= dx; tny -= dy; bny -= dy;
        int top = tly - tny, left = tlx - tnx;

        // copyMakeBorder(img, BORDER_REFLECT) then createLaplacePyr (8U branch: pyrDown / pyrUp on 8-bit images whose
        // values stay in [0,255], so the 16S arithmetic below gives the same numbers; subtract(..., CV_16S))
        Img16 bordered; bordered.rows = height; bordered.cols = width; bordered.d.resize((size_t)width * height * 3);
        for (int y = 0; y < height; y++) {
            int sy = border_reflect(y - top, ih, 0);
            for (int x = 0; x < width; x++) {
                int sx = border_reflect(x - left, iw, 0);
                const uint8_t* s = img + ((size_t)sy * iw + sx) * 3;
                int16_t* d = &bordered.d[((size_t)y * width + x) * 3];
                d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
            }
        }
        std::vector<Img16> src_lap;
        create_laplace_pyr(std::move(bordered), num_bands, src_lap);

        // weight map: mask.convertTo(CV_32F, 1/255.) or mask.convertTo(CV_16S) + 1 where mask != 0; copyMakeBorder CONSTANT 0
        std::vector<ImgF> wf(num_bands + 1);
        std::vector<Img16> ws(num_bands + 1);  // (cols used as single channel)
        if (blend == 2) {
            ws[0].rows = height; ws[0].cols = width; ws[0].d.assign((size_t)width * height, 0);
            for (int y = 0; y < ih; y++)
                for (int x = 0; x < iw; x++) {
                    int m = mask[(size_t)y * iw + x];
                    ws[0].d[(size_t)(y + top) * width + (x + left)] = (int16_t)(m + (m != 0));
                }
            for (int i = 0; i < num_bands; i++) {
                ws[i + 1].rows = (ws[i].rows + 1) / 2; ws[i + 1].cols = (ws[i].cols + 1) / 2;
                ws[i + 1].d.resize((size_t)ws[i + 1].rows * ws[i + 1].cols);
                pyr_down_s16(ws[i].d.data(), ws[i].rows, ws[i].cols, 1, ws[i + 1].d.data());
            }
        } else {
            wf[0].rows = height; wf[0].cols = width; wf[0].d.assign((size_t)width * height, 0.f);
            const float scale = (float)(1. / 255.);
            for (int y = 0; y < ih; y++)
                for (int x = 0; x < iw; x++) wf[0].d[(size_t)(y + top) * width + (x + left)] = (float)mask[(size_t)y * iw + x] * scale;
            for (int i = 0; i < num_bands; i++) {
                wf[i + 1].rows = (wf[i].rows + 1) / 2; wf[i + 1].cols = (wf[i].cols + 1) / 2;
                wf[i + 1].d.resize((size_t)wf[i + 1].rows * wf[i + 1].cols);
                pyr_down_f32(wf[i].d.data(), wf[i].rows, wf[i].cols, wf[i + 1].d.data());
            }
        }

        int y_tl = tny, y_br = bny, x_tl = tnx, x_br = bnx;
        for (int i = 0; i <= num_bands; ++i) {
            const int sc = src_lap[i].cols, dc = dst_lap[i].cols;
            for (int y = y_tl; y < y_br; ++y) {
                const int y_ = y - y_tl;
                const int16_t* src_row = &src_lap[i].d[(size_t)y_ * sc * 3];
                int16_t* dst_row = &dst_lap[i].d[(size_t)y * dc * 3];
                for (int x = x_tl; x < x_br; ++x) {
                    const int x_ = x - x_tl;
                    if (blend == 0) {
                        float w = wf[i].d[(size_t)y_ * sc + x_];
                        float& dw = dst_wf[i].d[(size_t)y * dc + x];
                        if (w >= dw) { dw = w; for (int c = 0; c < 3; c++) dst_row[3 * x + c] = src_row[3 * x_ + c]; }
                    } else if (blend == 1) {
                        float w = wf[i].d[(size_t)y_ * sc + x_];
                        for (int c = 0; c < 3; c++) dst_row[3 * x + c] = (int16_t)(dst_row[3 * x + c] + (int16_t)(int)(src_row[3 * x_ + c] * w));
                        dst_wf[i].d[(size_t)y * dc + x] += w;
                    } else {
                        int w = ws[i].d[(size_t)y_ * sc + x_];
                        for (int c = 0; c < 3; c++) dst_row[3 * x + c] = (int16_t)(dst_row[3 * x + c] + (int16_t)((src_row[3 * x_ + c] * w) >> 8));
                        int16_t& dw = dst_ws[i][(size_t)y * dc + x];
                        dw = (int16_t)(dw + w);
                    }
                }
            }
            x_tl /= 2; y_tl /= 2; x_br /= 2; y_br /= 2;
        }
    }

    // :256-300 + cv::detail::Blender::blend.  normalizeUsingWeightMap (OpenCV, stitching/blenders.cpp) runs only for the
    // weighted-sum blends: the reference's `should_normalize` is an uninitialised member that only its disabled branch sets.
    void finish(std::vector<int16_t>& result, std::vector<uint8_t>& mask) {
        if (blend != 0)
            for (int i = 0; i <= num_bands; ++i) {
                size_t n = (size_t)dst_lap[i].rows * dst_lap[i].cols;
                int16_t* p = dst_lap[i].d.data();
                if (blend == 1) {
                    const float eps = 1e-5f;
                    for (size_t k = 0; k < n; k++) {
                        float d = dst_wf[i].d[k] + eps;
                        for (int c = 0; c < 3; c++) p[3 * k + c] = (int16_t)(int)(p[3 * k + c] / d);
                    }
                } else {
                    for (size_t k = 0; k < n; k++) {
                        int w = dst_ws[i][k] + 1;
                        for (int c = 0; c < 3; c++) p[3 * k + c] = (int16_t)(w ? (p[3 * k + c] * 256) / w : 0);
                    }
                }
            }
        restore_from_laplace_pyr(dst_lap);
        result.assign((size_t)Wf * Hf * 3, 0);
        mask.assign((size_t)Wf * Hf, 0);
        const float epsf = (float)1e-5;
        for (int y = 0; y < Hf; y++)
            for (int x = 0; x < Wf; x++) {
                size_t s = (size_t)y * W + x, d = (size_t)y * Wf + x;
                bool on = (blend == 2) ? dst_ws[0][s] > 0 : dst_wf[0].d[s] > epsf;
                mask[d] = on ? 255 : 0;
                if (on) for (int c = 0; c < 3; c++) result[3 * d + c] = dst_lap[0].d[3 * s + c];
            }
    }
};

// Map2DRender::renderFrames, Map2DRender.cpp:479-760.  poses are camera-to-world; Map2DRender::feed / Map2DRenderPrepare::prepare
// move them to the plane frame first (:341-345, :443).
int render_frames(Map& m, RenderState& rs, int n, const uint8_t* const* frames, int fw, int fh, size_t stride, const double* poses7,
                  int blend, int bands_override, int* result) {
    rs = RenderState();
    if (!m.valid) return M2D_ERR_STATE;
    if (fw != m.cam_w || fh != m.cam_h) return M2D_REJECTED;
    std::vector<uint8_t> wimg((size_t)fw * fh);
    render_weight_image_u8(fw, fh, wimg.data());
    rs.frames.assign(n, RenderFrame());
    const double ipx[4] = {0, m.cam_w, 0, m.cam_w}, ipy[4] = {0, 0, m.cam_h, m.cam_h};
    double gminx = 0, gminy = 0, gmaxx = 0, gmaxy = 0;  // pi::Point2d min,max: default-constructed to (0,0), :532
    Pose pinv = pose_inverse(m.plane);
    for (int idx = 0; idx < n; idx++) {
        if (result) result[idx] = M2D_REJECTED;
        RenderFrame& F = rs.frames[idx];
        Pose f = pose_mul(pinv, pose_from7(poses7 + 7 * (size_t)idx));
        double px[4], py[4];
        Vec3 down{0, 0, -1};
        if (f.t.z < 0) down = Vec3{0, 0, 1};
        bool ok = true;
        for (int j = 0; j < 4; j++) {
            Vec3 axis = qrot(f.r, m.unproject(ipx[j], ipy[j]));
            if (axis.x * down.x + axis.y * down.y + axis.z * down.z < 0.4) { ok = false; break; }
            double s = f.t.z / axis.z;
            px[j] = f.t.x - axis.x * s; py[j] = f.t.y - axis.y * s;
        }
        if (!ok) continue;
        double cminx = 1e6, cminy = 1e6, cmaxx = -1e6, cmaxy = -1e6;
        for (int i = 0; i < 4; i++) {
            if (px[i] < cminx) cminx = px[i];
            if (py[i] < cminy) cminy = py[i];
            if (px[i] > cmaxx) cmaxx = px[i];
            if (py[i] > cmaxy) cmaxy = py[i];
        }
        if (cminx < gminx) gminx = cminx;
        if (cminy < gminy) gminy = cminy;
        if (cmaxx > gmaxx) gmaxx = cmaxx;
        if (cmaxy > gmaxy) gmaxy = cmaxy;
        F.cwx = (float)cminx; F.cwy = (float)cminy;
        F.iw = (int)((cmaxx - cminx) * m.length_pixel_inv);
        F.ih = (int)((cmaxy - cminy) * m.length_pixel_inv);
        if (F.iw <= 0 || F.ih <= 0) continue;  // cv::warpPerspective would throw on an empty size
        float srcp[8], dstp[8];
        for (int i = 0; i < 4; i++) {
            srcp[2 * i] = (float)ipx[i]; srcp[2 * i + 1] = (float)ipy[i];
            dstp[2 * i] = (float)((px[i] - cminx) * m.length_pixel_inv);
            dstp[2 * i + 1] = (float)((py[i] - cminy) * m.length_pixel_inv);
        }
        double M[9], Mi[9];
        if (!get_perspective_transform(srcp, dstp, M) || !invert3x3(M, Mi)) continue;
        F.img.resize((size_t)F.iw * F.ih * 3);
        F.mask.resize((size_t)F.iw * F.ih);
        warp_u8c3_linear_reflect(frames[idx], fh, fw, stride, Mi, F.img.data(), F.ih, F.iw);
        warp_u8c1_nearest_const0(wimg.data(), fh, fw, Mi, F.mask.data(), F.ih, F.iw);
        F.ok = true;
        if (result) result[idx] = M2D_OK;
    }
    // 2. spread the map (:606-621) and snap the area to whole tiles (:623-639)
    if (gminx < m.vmin.x || gminy < m.vmin.y || gmaxx > m.vmax.x || gmaxy > m.vmax.y)
        if (!m.spread(gminx, gminy, gmaxx, gmaxy)) return M2D_REJECTED;
    int xminInt = (int)std::floor((gminx - m.vmin.x) * m.ele_size_inv), yminInt = (int)std::floor((gminy - m.vmin.y) * m.ele_size_inv);
    int xmaxInt = (int)std::ceil((gmaxx - m.vmin.x) * m.ele_size_inv), ymaxInt = (int)std::ceil((gmaxy - m.vmin.y) * m.ele_size_inv);
    if (xminInt < 0 || yminInt < 0 || xmaxInt > m.w || ymaxInt > m.h || xminInt >= xmaxInt || yminInt >= ymaxInt) return M2D_REJECTED;
    double minx = m.vmin.x + m.ele_size * xminInt, miny = m.vmin.y + m.ele_size * yminInt;
    for (RenderFrame& F : rs.frames) {
        if (!F.ok) continue;
        F.cx = (int)((F.cwx - minx) * m.length_pixel_inv);
        F.cy = (int)((F.cwy - miny) * m.length_pixel_inv);
    }
    // 3. blend (:697-741)
    RenderBlender B;
    B.blend = blend;
    const int W = (xmaxInt - xminInt) * M2D_ELE_PIXELS, H = (ymaxInt - yminInt) * M2D_ELE_PIXELS;
    {
        double blend_strength = 5;
        float blend_width = std::sqrt(static_cast<float>(W * H)) * blend_strength / 100.f;
        B.actual_num_bands = static_cast<int>(std::ceil(std::log(blend_width) / std::log(2.)) - 1.);
        if (bands_override > 0) B.actual_num_bands = bands_override;
    }
    B.prepare(W, H);
    for (const RenderFrame& F : rs.frames)
        if (F.ok) B.feed(F.img.data(), F.mask.data(), F.iw, F.ih, F.cx, F.cy);
    B.finish(rs.result, rs.mask);
    rs.W = W; rs.H = H; rs.num_bands = B.num_bands;
    rs.tile_x0 = xminInt + m.org_x; rs.tile_y0 = yminInt + m.org_y;
    rs.have = true;
    m.last_rect[0] = xminInt; m.last_rect[1] = yminInt; m.last_rect[2] = xmaxInt; m.last_rect[3] = ymaxInt;
    return M2D_OK;
}

}  // namespace

#else  // RENDER_ORACLE_PART == 2: the extern "C" surface (inside map2d_oracle.cpp's extern "C" block)

static RenderState& render_state_of(orc_map* o) { return o->rs; }

// stand-in for m2d_render_frames
int orc_render_frames(orc_map* o, int n, const uint8_t* base, size_t frame_stride, int w, int h, size_t stride, const double* poses, int* result) {
    if (o->m.type != M2D_TYPE_RENDER) return M2D_ERR_STATE;
    std::vector<const uint8_t*> ptrs(n);
    for (int i = 0; i < n; i++) ptrs[i] = base + (size_t)i * frame_stride;
    return render_frames(o->m, render_state_of(o), n, ptrs.data(), w, h, stride, poses, o->m.cfg.render_blend, o->m.cfg.render_bands, result);
}
// stand-in for m2d_render_get: the 16SC3 result, the mask and the geometry
int orc_render_get(orc_map* o, int16_t* result16, uint8_t* mask, int* w, int* h, int* num_bands, int* tile_x0, int* tile_y0) {
    RenderState& rs = render_state_of(o);
    if (!rs.have) return M2D_REJECTED;
    if (w) *w = rs.W;
    if (h) *h = rs.H;
    if (num_bands) *num_bands = rs.num_bands;
    if (tile_x0) *tile_x0 = rs.tile_x0;
    if (tile_y0) *tile_y0 = rs.tile_y0;
    if (result16) memcpy(result16, rs.result.data(), rs.result.size() * sizeof(int16_t));
    if (mask) memcpy(mask, rs.mask.data(), rs.mask.size());
    return M2D_OK;
}
// the warped image / mask / canvas corner of frame i (what the reference hands to blender->feed), for the cv2 pin
int orc_render_warped(orc_map* o, int i, uint8_t* img, uint8_t* mask, int* iw, int* ih, int* cx, int* cy) {
    RenderState& rs = render_state_of(o);
    if (!rs.have || i < 0 || i >= (int)rs.frames.size()) return M2D_ERR_ARG;
    const RenderFrame& F = rs.frames[i];
    if (!F.ok) return M2D_REJECTED;
    *iw = F.iw; *ih = F.ih; *cx = F.cx; *cy = F.cy;
    if (img) memcpy(img, F.img.data(), F.img.size());
    if (mask) memcpy(mask, F.mask.data(), F.mask.size());
    return M2D_OK;
}
int orc_warp_u8c3_reflect(const uint8_t* src, int sh, int sw, const double* M, uint8_t* dst, int dh, int dw) {
    double Mi[9]; if (!invert3x3(M, Mi)) return 1;
    warp_u8c3_linear_reflect(src, sh, sw, (size_t)sw * 3, Mi, dst, dh, dw); return 0;
}
int orc_warp_u8c1_nearest(const uint8_t* src, int sh, int sw, const double* M, uint8_t* dst, int dh, int dw) {
    double Mi[9]; if (!invert3x3(M, Mi)) return 1;
    warp_u8c1_nearest_const0(src, sh, sw, Mi, dst, dh, dw); return 0;
}
void orc_render_weight_image(int w, int h, uint8_t* out) { render_weight_image_u8(w, h, out); }

#endif
