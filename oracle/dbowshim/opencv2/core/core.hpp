// Stand-in for <opencv2/core/core.hpp> -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Just enough of cv:: for the reference's vendored DBoW2 (Thirdparty/DBoW2/DBoW2/{TemplatedVocabulary.h, FORB.cpp, ...}) to
// compile UNMODIFIED where it lies: a dense matrix with an element size (descriptors are 1 x 32 CV_8U rows) and inert
// FileStorage / FileNode classes for the YAML save / load members (virtual, so they are instantiated, but the pin only
// calls loadFromTextFile and transform).
#pragma once
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_8UC1 CV_8U

namespace cv {

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    Mat() { }
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type)
    {
        const size_t es = type == CV_8U ? 1 : 4;
        if (r == rows && c == cols && es == esz_ && buf_) return;
        rows = r; cols = c; esz_ = es;
        buf_ = std::shared_ptr<std::vector<uchar>>(new std::vector<uchar>((size_t)r * c * es));
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, (size_t)r * c * m.esz_); return m; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    void release() { rows = cols = 0; data = nullptr; buf_.reset(); }
    Mat clone() const
    {
        Mat m;
        m.rows = rows; m.cols = cols; m.esz_ = esz_;
        if (buf_) { m.buf_ = std::shared_ptr<std::vector<uchar>>(new std::vector<uchar>(*buf_)); m.data = m.buf_->data(); }
        return m;
    }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * cols * esz_); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * cols * esz_); }
    template <typename T> T& at(int y, int x) { return ptr<T>(y)[x]; }
    template <typename T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }

private:
    size_t esz_ = 1;
    std::shared_ptr<std::vector<uchar>> buf_;
};

// never opened: save(filename) / load(filename) throw "Could not open file" like the reference on a missing file
class FileNode {
public:
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator float() const { return 0.f; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) { }
    FileStorage(const std::string&, int) { }
    bool isOpened() const { return false; }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

} // namespace cv
