// Drives include/Map2DB200.h exactly like Map2DFusion/Map2DFusion.cpp:273-329 drives a Map2D:
// create -> prepare(plane, camera, frames) -> feed(img, pose)* -> save().  Prints one line the test parses.
#include <cstdio>
#include <cstring>
#include "Map2DB200.h"

int main(int argc, char** argv) {
    int type = argc > 1 ? atoi(argv[1]) : Map2D::TypeMultiBandCPU;
    const char* out = argc > 2 ? argv[2] : "/tmp/adapter_stub.png";
    const bool threaded = argc > 3 && atoi(argv[3]) != 0;   // Map2D::create(type, thread)
    const int n_feed = argc > 4 ? atoi(argv[4]) : 4;        // frames handed to feed() ...
    const int first_feed = argc > 5 ? atoi(argv[5]) : 0;    // ... starting at this one (frames 0..3 are the prepare set)
    const int n_dev = argc > 6 ? atoi(argv[6]) : 1;         // > 1: one process, several GPUs (m2d_create_multi)
    const int W = 320, H = 180;
    std::vector<int> devices;
    for (int d = 0; d < n_dev; d++) devices.push_back(d);
    Map2DB200 map(type, threaded, devices);
    std::deque<std::pair<cv::Mat, pi::SE3d> > frames, all;
    for (int k = 0; k < 8; k++) {
        cv::Mat img(H, W, CV_8UC3);
        for (size_t i = 0; i < img.buf.size(); i++) img.buf[i] = (unsigned char)((i * 7 + k * 31) & 255);
        all.push_back(std::make_pair(img, pi::SE3d(10.0 * k, 5.0 * k, 100, 1, 0, 0, 0)));  // nadir: 180 deg about X
        if (k < 4) frames.push_back(all.back());
    }
    pi::SE3d plane;  // identity
    PinHoleParameters cam(W, H, 0.9 * W, 0.9 * W, W / 2.0, H / 2.0);
    bool prepared = map.prepare(plane, cam, frames);
    int fed = 0;
    for (int k = first_feed; prepared && k < first_feed + n_feed && k < (int)all.size(); k++) fed += map.feed(all[k].first, all[k].second) ? 1 : 0;
    bool oblique = prepared && map.feed(frames[0].first, pi::SE3d(0, 0, 100, 0.5, 0.5, 0.5, 0.5));
    bool saved = prepared && map.save(out);
    cv::Mat img = prepared ? map.getImage() : cv::Mat();
    printf("handle=%d prepared=%d fed=%d oblique_accepted=%d saved=%d image=%dx%d queue=%u\n", map.handle() != NULL, prepared, fed,
           oblique, saved, img.cols, img.rows, map.queueSize());
    return 0;
}
