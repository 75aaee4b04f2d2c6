// mini_cv.h -- TEST INFRASTRUCTURE (oracle/): the part of OpenCV's C++ API that the reference's hot-path sources use, so that
// they compile VERBATIM into oracle/_ref/ without OpenCV headers (not in this image).  The four imgproc
// kernels the path calls (cvtColor, pyrDown, Canny, distanceTransform) are served by the C restatements of
// oracle/revo_oracle.c, which tests/test_oracle.py pins bit for bit against python cv2 4.13; everything else is plain
// container code.  Debug-only calls of the reference (imshow, imwrite, waitKey, merge, normalize, remap, GaussianBlur) are
// no-ops or aborts: the live path never reaches them (imwrite: the reference writes debug PNGs, nothing reads them).
#pragma once
#include <sys/types.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

extern "C" {
void orc_gray_bgr(const uint8_t *bgr, int w, int h, size_t stride, int ch, uint8_t *gray);
void orc_pyrdown_u8(const uint8_t *src, int w, int h, uint8_t *dst);
void orc_canny(const uint8_t *gray, int w, int h, double t1, double t2, uint8_t *out);
void orc_edt_l2(const uint8_t *edges, int w, int h, float *dt);
}

#define CV_8U 0
#define CV_16S 3
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_BGRA2GRAY 10
#define CV_INTER_LINEAR 1
#define CV_DIST_L2 2
#define CV_DIST_MASK_PRECISE 0

namespace cv {

struct Size2i {
    int width = 0, height = 0;
    Size2i() {}
    Size2i(int w, int h) : width(w), height(h) {}
};
typedef Size2i Size;
inline std::ostream &operator<<(std::ostream &o, const Size2i &s) { return o << "[" << s.width << " x " << s.height << "]"; }
struct Rect { int x, y, width, height; Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
struct Scalar { double v[4]; Scalar(double a = 0) { v[0] = a; v[1] = v[2] = v[3] = 0; } };
struct Vec3b {
    uint8_t v[3];
    Vec3b() { v[0] = v[1] = v[2] = 0; }
    Vec3b(uint8_t a, uint8_t b, uint8_t c) { v[0] = a; v[1] = b; v[2] = c; }
    uint8_t operator[](int i) const { return v[i]; }
};
enum { NORM_MINMAX = 32 };

class Mat {
public:
    int rows = 0, cols = 0, type_ = 0;
    uint8_t *data = nullptr;
    std::shared_ptr<std::vector<uint8_t>> buf;      // reference-counted pixels: copies are shallow like cv::Mat's

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &s) { create(r, c, type); setTo(s); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        buf = std::make_shared<std::vector<uint8_t>>((size_t)r * c * elemSize() + 64, 0);
        data = buf->data();
    }
    static int depthSize(int type) { const int d = type & 7; return d == CV_8U ? 1 : d == CV_16S ? 2 : d == CV_32F ? 4 : 8; }
    int channels() const { return (type_ >> 3) + 1; }
    int type() const { return type_; }
    size_t elemSize() const { return (size_t)depthSize(type_) * channels(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || total() == 0; }
    Size2i size() const { return Size2i(cols, rows); }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; }
    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        std::memcpy(m.data, data, total() * elemSize());
        return m;
    }
    void copyTo(Mat &dst) const { dst = clone(); }
    Mat &setTo(const Scalar &s)
    {
        const int d = type_ & 7;
        for (size_t i = 0; i < total() * channels(); ++i) {
            if (d == CV_8U) data[i] = (uint8_t)s.v[0];
            else if (d == CV_32F) ((float *)data)[i] = (float)s.v[0];
            else if (d == CV_64F) ((double *)data)[i] = s.v[0];
        }
        return *this;
    }
    template <typename T> T &at(int y, int x) { return ((T *)data)[(size_t)y * cols + x]; }
    template <typename T> const T &at(int y, int x) const { return ((const T *)data)[(size_t)y * cols + x]; }
    static Mat eye(int r, int c, int type)
    {
        Mat m(r, c, type, Scalar(0));
        for (int i = 0; i < r && i < c; ++i) m.at<double>(i, i) = 1.0;
        return m;
    }
    void convertTo(Mat &dst, int type, double alpha = 1.0) const
    {
        Mat o(rows, cols, type);
        for (size_t i = 0; i < total(); ++i) {
            double v = (type_ & 7) == CV_32F ? ((const float *)data)[i] : (type_ & 7) == CV_8U ? data[i] : ((const double *)data)[i];
            v *= alpha;
            if ((type & 7) == CV_8U) o.data[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : std::lrint(v));
            else if ((type & 7) == CV_32F) ((float *)o.data)[i] = (float)v;
            else ((double *)o.data)[i] = v;
        }
        dst = o;
    }
    Mat operator()(const Rect &) const { return *this; }      // debug printing of a region only
};
inline std::ostream &operator<<(std::ostream &o, const Mat &m) { return o << "Mat(" << m.rows << "x" << m.cols << ")"; }

// MatExpr stand-ins (eager): 255 - m (8U, saturating), m / s, m * s (float images, debug output only)
inline Mat operator-(int a, const Mat &m)
{
    Mat o(m.rows, m.cols, m.type_);
    for (size_t i = 0; i < m.total() * m.channels(); ++i) { const int v = a - (int)m.data[i]; o.data[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
    return o;
}
inline Mat operator/(const Mat &m, double s) { Mat o; m.convertTo(o, m.type_, 1.0 / s); return o; }
inline Mat operator*(const Mat &m, double s) { Mat o; m.convertTo(o, m.type_, s); return o; }

struct Matx33d { static Matx33d eye() { return Matx33d(); } };

// ---- imgproc: the four kernels of the path ---------------------------------------------------------------------
inline void cvtColor(const Mat &src, Mat &dst, int code)
{
    if (code != CV_BGRA2GRAY || (src.channels() != 3 && src.channels() != 4)) { std::fprintf(stderr, "mini_cv: unsupported cvtColor\n"); std::abort(); }
    Mat o(src.rows, src.cols, CV_8UC1);
    orc_gray_bgr(src.data, src.cols, src.rows, (size_t)src.cols * src.channels(), src.channels(), o.data);
    dst = o;
}
inline void pyrDown(const Mat &src, Mat &dst)
{
    const int w = src.cols, h = src.rows, wd = (w + 1) / 2, hd = (h + 1) / 2, ch = src.channels();
    Mat o(hd, wd, src.type_);
    if (ch == 1) {
        orc_pyrdown_u8(src.data, w, h, o.data);
    } else {      // per channel (generateColoredPcl down-sizes the colour image)
        std::vector<uint8_t> a((size_t)w * h), b((size_t)wd * hd);
        for (int c = 0; c < ch; ++c) {
            for (size_t i = 0; i < (size_t)w * h; ++i) a[i] = src.data[i * ch + c];
            orc_pyrdown_u8(a.data(), w, h, b.data());
            for (size_t i = 0; i < (size_t)wd * hd; ++i) o.data[i * ch + c] = b[i];
        }
    }
    dst = o;
}
inline void Canny(const Mat &gray, Mat &edges, double t1, double t2, int aperture, bool l2)
{
    if (aperture != 3 || !l2 || gray.channels() != 1) { std::fprintf(stderr, "mini_cv: unsupported Canny\n"); std::abort(); }
    Mat o(gray.rows, gray.cols, CV_8UC1);
    orc_canny(gray.data, gray.cols, gray.rows, t1, t2, o.data);
    edges = o;
}
inline void distanceTransform(const Mat &src, Mat &dst, int dist_type, int mask)
{
    if (dist_type != CV_DIST_L2 || mask != CV_DIST_MASK_PRECISE) { std::fprintf(stderr, "mini_cv: unsupported distanceTransform\n"); std::abort(); }
    // distance to the nearest ZERO pixel of src
    std::vector<uint8_t> feat(src.total());
    for (size_t i = 0; i < src.total(); ++i) feat[i] = src.data[i] == 0 ? 255 : 0;
    Mat o(src.rows, src.cols, CV_32FC1);
    orc_edt_l2(feat.data(), src.cols, src.rows, (float *)o.data);
    dst = o;
}
inline int countNonZero(const Mat &m)
{
    int n = 0;
    for (size_t i = 0; i < m.total(); ++i) n += m.data[i] != 0;
    return n;
}
inline void add(const Mat &a, const Mat &b, Mat &c)       // 8U, saturating
{
    Mat o(a.rows, a.cols, a.type_);
    for (size_t i = 0; i < a.total(); ++i) { const int v = (int)a.data[i] + (int)b.data[i]; o.data[i] = (uint8_t)(v > 255 ? 255 : v); }
    c = o;
}
inline void minMaxIdx(const Mat &m, double *mn, double *mx)
{
    double lo = 1e300, hi = -1e300;
    for (size_t i = 0; i < m.total(); ++i) {
        const double v = (m.type_ & 7) == CV_32F ? ((const float *)m.data)[i] : m.data[i];
        if (v < lo) lo = v;
        if (v > hi) hi = v;
    }
    if (mn) *mn = lo;
    if (mx) *mx = hi;
}
// debug / dead code of the reference: never reached on the live path
inline void remap(const Mat &, Mat &, const Mat &, const Mat &, int) { std::fprintf(stderr, "mini_cv: remap is not available\n"); std::abort(); }
inline void GaussianBlur(const Mat &, Mat &, Size, double) { std::fprintf(stderr, "mini_cv: GaussianBlur is not available\n"); std::abort(); }
inline void merge(const std::vector<Mat> &, Mat &) { std::fprintf(stderr, "mini_cv: merge is not available\n"); std::abort(); }
inline void normalize(const Mat &, Mat &, double, double, int) { std::fprintf(stderr, "mini_cv: normalize is not available\n"); std::abort(); }
inline Mat getOptimalNewCameraMatrix(const Mat &, const Mat &, Size, double, Size) { std::abort(); }
inline void initUndistortRectifyMap(const Mat &, const Mat &, const Matx33d &, const Mat &, Size, int, Mat &, Mat &) { std::abort(); }
inline void imshow(const std::string &, const Mat &) {}
inline bool imwrite(const std::string &, const Mat &) { return true; }
inline int waitKey(int = 0) { return 0; }

// ---- FileStorage: "no file": every cv::read takes its default (the harness then sets the fields it wants) ------------
struct FileNode {};
class FileStorage {
public:
    enum { READ = 0 };
    FileStorage(const std::string &, int) {}
    bool isOpened() const { return true; }
    FileNode operator[](const char *) const { return FileNode(); }
    void release() {}
};
template <typename T, typename D>
inline void read(const FileNode &, T &v, const D &def) { v = (T)def; }

}  // namespace cv
