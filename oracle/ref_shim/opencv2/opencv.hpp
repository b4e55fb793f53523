// TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for <opencv2/opencv.hpp>.
//
// The reference (RunhuiWang/DeltaPQ) includes OpenCV in pq.h:8 and utils.h:4 but the
// query / encode / tree-build paths only need the `uchar` typedef plus the std headers
// OpenCV drags in.  pq.cpp additionally needs a tiny cv::Mat (pq.cpp:125-151,314-339)
// and cv::kmeans (pq.cpp:149, pq_tree.cpp:103), which only the out-of-scope `learn`
// task reaches; here kmeans aborts.  OpenCV C++ headers are not installed in this
// image, so this file is put on the include path by oracle/Makefile when it compiles
// the UNMODIFIED reference sources from /root/reference.
#ifndef DPQ_ORACLE_OPENCV_SHIM_HPP
#define DPQ_ORACLE_OPENCV_SHIM_HPP

#include <algorithm>
#include <array>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_32FC1 5

namespace cv {

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT_MIN, INT_MAX); }
};

// Row-major float matrix sharing storage between views (enough for pq.cpp).
class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), type_(CV_32FC1), stride_(0), off_(0) {}
    Mat(int r, int c, int type)
        : rows(r), cols(c), type_(type), stride_(c), off_(0),
          buf_(new std::vector<float>((size_t)r * c, 0.f)) {}
    int type() const { return type_; }
    template <typename T> T& at(int r, int c) {
        return reinterpret_cast<T&>((*buf_)[off_ + (size_t)r * stride_ + c]);
    }
    template <typename T> const T& at(int r, int c) const {
        return reinterpret_cast<const T&>((*buf_)[off_ + (size_t)r * stride_ + c]);
    }
    Mat operator()(const Range& rr, const Range& cr) const {
        Mat m = *this;
        int r0 = rr.start == INT_MIN ? 0 : rr.start, r1 = rr.end == INT_MAX ? rows : rr.end;
        int c0 = cr.start == INT_MIN ? 0 : cr.start, c1 = cr.end == INT_MAX ? cols : cr.end;
        m.off_ = off_ + (size_t)r0 * stride_ + c0;
        m.rows = r1 - r0;
        m.cols = c1 - c0;
        return m;
    }
private:
    int type_;
    size_t stride_, off_;
    std::shared_ptr<std::vector<float> > buf_;
};

struct TermCriteria {
    enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria(int t, int n, double e) : type(t), maxCount(n), epsilon(e) {}
};

enum { KMEANS_RANDOM_CENTERS = 0, KMEANS_PP_CENTERS = 2 };

inline double kmeans(const Mat&, int, Mat&, TermCriteria, int, int, Mat&) {
    std::fprintf(stderr, "opencv shim: cv::kmeans is not available in the oracle build "
                         "(the `learn` task is outside the parity-pinned path)\n");
    std::abort();
    return 0.0;
}

}  // namespace cv

#endif
