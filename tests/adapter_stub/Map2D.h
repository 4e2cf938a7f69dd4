// Minimal stand-ins for the reference types Map2DB200.h touches (Map2DFusion/Map2D.h, cv::Mat, pi::SE3d, svar), so
// that the adapter can be compiled and exercised in this repository, which has no OpenCV C++/Qt/GL headers.
// NOT the reference's code: only the member names/signatures the adapter uses are declared.
#pragma once
#include <deque>
#include <map>
#include <string>
#include <utility>
#include <vector>
typedef unsigned int uint;
enum { CV_8UC3 = 16, CV_8UC4 = 24 };
namespace cv {
struct Mat {
    int rows, cols, _type;
    size_t step;
    unsigned char* data;
    std::vector<unsigned char> buf;
    Mat() : rows(0), cols(0), _type(CV_8UC3), step(0), data(0) {}
    Mat(int r, int c, int t) : rows(r), cols(c), _type(t), step((size_t)c * (t == CV_8UC4 ? 4 : 3)), buf((size_t)r * c * (t == CV_8UC4 ? 4 : 3)) { data = buf.data(); }
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), _type(o._type), step(o.step), buf(o.buf) { data = buf.empty() ? o.data : buf.data(); }
    int type() const { return _type; }
    bool empty() const { return !data; }
};
}
namespace pi {
struct Point3d { double x, y, z; };
struct SO3d { double x, y, z, w; };
struct SE3d {
    SO3d r; Point3d t;
    SE3d() { r.x = r.y = r.z = 0; r.w = 1; t.x = t.y = t.z = 0; }
    SE3d(double X, double Y, double Z, double qx, double qy, double qz, double qw) { t.x = X; t.y = Y; t.z = Z; r.x = qx; r.y = qy; r.z = qz; r.w = qw; }
    const Point3d& get_translation() const { return t; }
    const SO3d& get_rotation() const { return r; }
};
}
struct SvarStub {
    std::map<std::string, double> d;
    double GetDouble(const std::string& k, double def = 0) { return d.count(k) ? d[k] : def; }
    int GetInt(const std::string& k, int def = 0) { return d.count(k) ? (int)d[k] : def; }
};
static SvarStub svar;
struct PinHoleParameters {
    PinHoleParameters() {}
    PinHoleParameters(int _w, int _h, double _fx, double _fy, double _cx, double _cy) : w(_w), h(_h), fx(_fx), fy(_fy), cx(_cx), cy(_cy) {}
    double w, h, fx, fy, cx, cy;
};
class Map2D {
public:
    enum Map2DType { NoType = 0, TypeCPU = 1, TypeGPU = 2, TypeMultiBandCPU = 3, TypeRender = 4 };
    virtual ~Map2D() {}
    virtual bool prepare(const pi::SE3d&, const PinHoleParameters&, const std::deque<std::pair<cv::Mat, pi::SE3d> >&) { return false; }
    virtual bool feed(cv::Mat, const pi::SE3d&) { return false; }
    virtual void draw() {}
    virtual bool save(const std::string&) { return false; }
    virtual uint queueSize() { return 0; }
};
