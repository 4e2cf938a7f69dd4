/*
 * Map2DB200.h — header-only C++ adapter: `class Map2DB200 : public Map2D` over the C-ABI of map2d_b200.h.
 *
 * This is the binding a pi-slam-fusion maintainer adds to the reference tree (next to Map2DFusion/Map2DCPU.h) so
 * that `Map2D::create(type)` hands the existing host code (Map2DFusion/Map2DFusion.cpp:273-329, the only caller)
 * a B200-backed object with unchanged semantics.  It is compiled in the REFERENCE build (it needs the reference's
 * Map2D.h, cv::Mat and pi::SE3d); this repository cannot compile it (no OpenCV C++/Qt/GL headers here) and ships
 * tests/adapter_stub/ to syntax-check it against minimal stand-ins.  See INTEGRATION.md.
 *
 *   reference member                         -> C-ABI call
 *   Map2D::prepare(plane,camera,frames)      -> m2d_prepare        (Map2D.h:88,  Map2DCPU.cpp:105-125)
 *   Map2D::feed(img,pose)                    -> m2d_feed           (Map2D.h:91,  Map2DCPU.cpp:127-148)
 *   Map2D::save(filename)                    -> m2d_save           (Map2D.h:95,  Map2DCPU.cpp:523-564)
 *   Map2D::queueSize()                       -> m2d_queue_size     (Map2D.h:97)
 *   thread=true (Map2DCPU.cpp:119-120,139-142,397-413: worker thread + frame queue of 20, drop-oldest; the queue
 *   starts out holding the prepare-frames, Map2D.cpp:42, which the worker therefore renders first)
 *                                            -> m2d_ingest_open_seeded(20, n) + m2d_ingest_push of every prepare-frame
 *                                               at prepare(); feed() = m2d_ingest_push
 *   Map2D::TypeRender (Map2DRender.cpp)      -> prepare(thread=true) hands the prepare-frames to m2d_render_frames as the ONE batch
 *                                               the reference's worker renders before it stops (:419-420, 758, 812-829) and
 *                                               writes "result.png" (reference: "result.jpg", :748); feed() only queues
 *                                               (thread=true, :446-452) or returns false (thread=false, :464-467)
 *   Map2D::draw()                            -> no-op (GL is out of scope); its data path is m2d_poll_changed +
 *                                               m2d_get_tile_image (changed tiles, blended textures) and
 *                                               m2d_tile_gps_corners (the Map2DUpdate overlay corners)
 */
#ifndef MAP2D_B200_ADAPTER_H
#define MAP2D_B200_ADAPTER_H

#include <algorithm>
#include <deque>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include "Map2D.h"       /* reference: Map2DFusion/Map2D.h (Map2D, PinHoleParameters, cv::Mat, pi::SE3d, svar) */
#include "map2d_b200.h"

class Map2DB200 : public Map2D {
public:
    /* type: Map2D::TypeCPU / TypeGPU (weighted), TypeMultiBandCPU or TypeRender.  The svar keys the CPU classes read
     * (Map2DCPU.cpp:75,246; MultiBandMap2DCPU.cpp:228,235,260,444,840) are copied into the config once. */
    explicit Map2DB200(int type, bool thread = true, int device = 0) : _h(NULL), _thread(thread), _type(type) {
        std::vector<int> devices(1, device);
        init(type, thread, devices);
    }
    /* Several GPUs behind the same object (one host process, tiles sharded over the devices): m2d_create_multi. */
    Map2DB200(int type, bool thread, const std::vector<int>& devices) : _h(NULL), _thread(thread), _type(type) { init(type, thread, devices); }
    virtual ~Map2DB200() { m2d_destroy(_h); }

private:
    void init(int type, bool thread, const std::vector<int>& devices) {
        m2d_config cfg;
        m2d_config_default(&cfg);
        cfg.scale = svar.GetDouble("Map2D.Scale", 1);
        cfg.resolution = svar.GetDouble("Map2D.Resolution", 0);
        cfg.weight_type = svar.GetInt("Map2D.WeightType", 0);
        cfg.band_number = svar.GetInt("MultiBandMap2DCPU.BandNumber", 5);
        cfg.force_float = svar.GetInt("MultiBandMap2DCPU.ForceFloat", 0);
        cfg.background = svar.GetInt("Result.BackGroundColor");
        cfg.thread = thread ? 1 : 0;
        cfg.device = devices.empty() ? 0 : devices[0];
        int rc = devices.size() > 1 ? m2d_create_multi(type, &cfg, (int)devices.size(), &devices[0], &_h) : m2d_create(type, &cfg, &_h);
        if (rc != M2D_OK) std::cerr << "Map2DB200: m2d_create failed (" << rc << "); no CPU fallback is taken.\n";
    }

public:

    virtual bool prepare(const pi::SE3d& plane, const PinHoleParameters& camera,
                         const std::deque<std::pair<cv::Mat, pi::SE3d> >& frames) {
        if (!_h) return false;
        double p[7], cam[6] = {camera.w, camera.h, camera.fx, camera.fy, camera.cx, camera.cy};
        pose7(plane, p);
        std::vector<double> poses(frames.size() * 7);
        size_t i = 0;
        for (std::deque<std::pair<cv::Mat, pi::SE3d> >::const_iterator it = frames.begin(); it != frames.end(); ++it, ++i)
            pose7(it->second, &poses[7 * i]);
        /* A second prepare() swaps in a new Prepare object with its own frame deque (Map2DCPU.cpp:105-125): whatever
         * was still queued for the old map is never rendered.  So: discard the old queue BEFORE the grid is replaced. */
        if (_thread && _type != M2D_TYPE_RENDER) m2d_ingest_abort(_h);
        if (m2d_prepare(_h, p, cam, (int)frames.size(), poses.empty() ? NULL : &poses[0]) != M2D_OK) return false;
        if (_type == M2D_TYPE_RENDER) return _thread ? renderBatch(frames) : true;
        if (!_thread) return true;   /* thread=false: the prepare-frames are never rendered (only feed() renders) */
        /* thread=true: the reference starts its worker here (Map2DCPU.cpp:119-120).  Map2DPrepare::_frames is both the
         * prepare set and the worker's queue (Map2D.cpp:42-47), so the worker renders the prepare-frames first, in
         * order; feed() then appends and drops one oldest entry beyond 20 (Map2DCPU.cpp:139-142). */
        if (m2d_ingest_open_seeded(_h, 20, (int)frames.size(), 1) != M2D_OK) return false;
        bool ok = true;
        for (std::deque<std::pair<cv::Mat, pi::SE3d> >::const_iterator it = frames.begin(); it != frames.end(); ++it) {
            const cv::Mat& img = it->first;
            double q[7];
            pose7(it->second, q);   /* the original camera-to-world pose: the library applies plane^-1 itself */
            if (img.type() != CV_8UC3 || !img.data) continue;   /* renderFrame would reject it (Map2DCPU.cpp:158-162) */
            int rc = m2d_ingest_push(_h, img.data, img.cols, img.rows, img.step, 3, q);
            if (rc < 0) ok = false;
        }
        return m2d_ingest_pause(_h, 0) == M2D_OK && ok;
    }

    /* pose is camera-to-world; the library left-multiplies plane^-1 like Map2DCPU.cpp:136. */
    virtual bool feed(cv::Mat img, const pi::SE3d& pose) {
        if (!_h) return false;
        if (img.type() != CV_8UC3) {  /* Map2DCPU.cpp:158-162 */
            std::cerr << "Map2DB200::feed: frame.type()!=CV_8UC3\n";
            return false;
        }
        double p[7];
        pose7(pose, p);
        if (_type == M2D_TYPE_RENDER && _thread) return true;   /* queued, never rendered: the worker stopped after its first batch */
        if (_thread) return m2d_ingest_push(_h, img.data, img.cols, img.rows, img.step, 3, p) == M2D_OK;  /* enqueue only */
        return m2d_feed(_h, img.data, img.cols, img.rows, img.step, p) == M2D_OK;
    }

    virtual void draw() {}

    /* thread=true: frames still queued are fused first (the reference would save without them). */
    virtual bool save(const std::string& filename) {
        if (!_h) return false;
        if (_thread && _type != M2D_TYPE_RENDER) m2d_ingest_drain(_h);
        return m2d_save(_h, filename.c_str()) == M2D_OK;   /* (Map2DRender::save returns false; here the last canvas is written) */
    }

    virtual uint queueSize() { return _h ? (uint)m2d_queue_size(_h) : 0; }

    /* Addition over the reference: the saved image in memory, BGRA (weighted) or BGR. */
    cv::Mat getImage(int* tile_min_x = NULL, int* tile_min_y = NULL) {
        int w, h, cn, tx, ty;
        if (_h && _thread && _type != M2D_TYPE_RENDER) m2d_ingest_drain(_h);
        if (!_h || m2d_get_image(_h, NULL, &w, &h, &cn, &tx, &ty) != M2D_OK) return cv::Mat();
        cv::Mat out(h, w, cn == 4 ? CV_8UC4 : CV_8UC3);
        if (m2d_get_image(_h, out.data, &w, &h, &cn, &tx, &ty) != M2D_OK) return cv::Mat();
        if (tile_min_x) *tile_min_x = tx;
        if (tile_min_y) *tile_min_y = ty;
        return out;
    }

    m2d_handle handle() const { return _h; }

private:
    /* Map2DRender::run -> renderFrames(frames): everything queued -- at prepare() time, the prepare-frames -- is one batch. */
    bool renderBatch(const std::deque<std::pair<cv::Mat, pi::SE3d> >& frames) {
        if (frames.empty()) return true;
        const cv::Mat& f0 = frames.front().first;
        const size_t frame_bytes = (size_t)f0.cols * f0.rows * 3;
        std::vector<unsigned char> pix(frame_bytes * frames.size());
        std::vector<double> poses(frames.size() * 7);
        int n = 0;
        for (std::deque<std::pair<cv::Mat, pi::SE3d> >::const_iterator it = frames.begin(); it != frames.end(); ++it) {
            const cv::Mat& img = it->first;
            if (img.type() != CV_8UC3 || !img.data || img.cols != f0.cols || img.rows != f0.rows) continue;
            for (int y = 0; y < img.rows; y++)
                std::copy(img.data + (size_t)y * img.step, img.data + (size_t)y * img.step + (size_t)img.cols * 3,
                          pix.begin() + frame_bytes * n + (size_t)y * img.cols * 3);
            pose7(it->second, &poses[7 * (size_t)n]);
            n++;
        }
        if (!n || m2d_render_frames(_h, n, &pix[0], frame_bytes, f0.cols, f0.rows, (size_t)f0.cols * 3, &poses[0], 0, NULL) != M2D_OK) return false;
        m2d_save(_h, "result.png");   /* cv::imwrite("result.jpg", result), Map2DRender.cpp:748 */
        return true;
    }
    static void pose7(const pi::SE3d& s, double* o) { /* stream order x y z qx qy qz qw, SE3.h:105-117 */
        const pi::Point3d& t = s.get_translation();
        const pi::SO3d& r = s.get_rotation();
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = r.x; o[4] = r.y; o[5] = r.z; o[6] = r.w;
    }
    m2d_handle _h;
    bool _thread;
    int _type;
    Map2DB200(const Map2DB200&);
    Map2DB200& operator=(const Map2DB200&);
};

#endif /* MAP2D_B200_ADAPTER_H */
