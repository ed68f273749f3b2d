// Minimal stand-in for the four OpenCV calls the reference's CIFAR dataset makes
// (/root/reference/dcifar10/common/custom.hpp:35-46: imread, Mat::empty, resize, split, Mat::ptr) so the
// UNMODIFIED reference compiles on an image without OpenCV C++.  There are no dataset files offline and the
// reference's path is a hard-coded AFS location, so imread() SYNTHESISES a deterministic 32x32 BGR image from
// the path string (cheaper than a real JPEG decode => this stub can only flatter the reference's timing).
// If EGCV_ROOT is set and <EGCV_ROOT>/<relative path>.raw exists (HxWx3 uint8, header "EGCV h w\n") it is read
// instead.  Not part of the product.
#ifndef EGCV_OPENCV_HPP
#define EGCV_OPENCV_HPP
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace cv {

struct Size {
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};

class Mat {
  public:
    int rows = 0, cols = 0, chans = 0;
    std::shared_ptr<std::vector<unsigned char>> buf;
    Mat() {}
    Mat(int r, int c, int ch) : rows(r), cols(c), chans(ch), buf(std::make_shared<std::vector<unsigned char>>((size_t)r * c * ch)) {}
    bool empty() const { return !buf || buf->empty(); }
    int channels() const { return chans; }
    unsigned char *ptr(int row = 0) { return buf->data() + (size_t)row * cols * chans; }
    const unsigned char *ptr(int row = 0) const { return buf->data() + (size_t)row * cols * chans; }
};

inline Mat imread(const std::string &path, int /*flags*/ = 1) {
    if (const char *root = std::getenv("EGCV_ROOT")) {
        std::string p = std::string(root) + "/" + path + ".raw";
        if (FILE *f = std::fopen(p.c_str(), "rb")) {
            int h = 0, w = 0;
            Mat m;
            if (std::fscanf(f, "EGCV %d %d\n", &h, &w) == 2 && h > 0 && w > 0) {
                m = Mat(h, w, 3);
                if (std::fread(m.ptr(), 1, (size_t)h * w * 3, f) != (size_t)h * w * 3) m = Mat();
            }
            std::fclose(f);
            if (!m.empty()) return m;
        }
    }
    // FNV-1a of the path seeds an xorshift stream; class-dependent mean so the data is learnable
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : path) { h ^= c; h *= 1099511628211ull; }
    Mat m(32, 32, 3);
    unsigned char *d = m.ptr();
    uint64_t s = h | 1;
    unsigned bias = (unsigned)(h >> 56) & 0x3f;
    for (int i = 0; i < 32 * 32 * 3; i += 8) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        for (int b = 0; b < 8; b++) d[i + b] = (unsigned char)((((s >> (8 * b)) & 0xff) * 3 >> 2) + bias);
    }
    return m;
}

inline void resize(const Mat &src, Mat &dst, Size sz, double = 0, double = 0, int = 1) {
    if (src.rows == sz.height && src.cols == sz.width) {
        if (&src != &dst) dst = src;
        return;
    }
    Mat out(sz.height, sz.width, src.chans);  // bilinear, pixel-centre aligned like INTER_LINEAR
    const float sy = (float)src.rows / sz.height, sx = (float)src.cols / sz.width;
    for (int y = 0; y < sz.height; y++) {
        float fy = (y + 0.5f) * sy - 0.5f; if (fy < 0) fy = 0;
        int y0 = (int)fy, y1 = y0 + 1 < src.rows ? y0 + 1 : y0; float wy = fy - y0;
        for (int x = 0; x < sz.width; x++) {
            float fx = (x + 0.5f) * sx - 0.5f; if (fx < 0) fx = 0;
            int x0 = (int)fx, x1 = x0 + 1 < src.cols ? x0 + 1 : x0; float wx = fx - x0;
            for (int c = 0; c < src.chans; c++) {
                float v = (1 - wy) * ((1 - wx) * src.ptr(y0)[x0 * src.chans + c] + wx * src.ptr(y0)[x1 * src.chans + c]) +
                          wy * ((1 - wx) * src.ptr(y1)[x0 * src.chans + c] + wx * src.ptr(y1)[x1 * src.chans + c]);
                out.ptr(y)[x * src.chans + c] = (unsigned char)(v + 0.5f);
            }
        }
    }
    dst = out;
}

inline void split(const Mat &src, std::vector<Mat> &planes) {
    planes.resize(src.chans);
    for (int c = 0; c < src.chans; c++) {
        planes[c] = Mat(src.rows, src.cols, 1);
        unsigned char *o = planes[c].ptr();
        const unsigned char *s = src.ptr();
        for (int i = 0; i < src.rows * src.cols; i++) o[i] = s[i * src.chans + c];
    }
}

}  // namespace cv
#endif
