// oracle/minicv/dbow_shim.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Force-included (-include) in front of the reference's vendored Thirdparty/DBoW2 sources so that they compile UNMODIFIED
// against oracle/minicv: the stream headers that real OpenCV pulls in transitively, and inert cv::FileStorage / cv::FileNode
// stand-ins for the YAML save()/load() members (virtual, so their bodies must compile; the oracle only ever uses
// loadFromTextFile + transform, and the stand-ins abort if they are reached).
#ifndef ORACLE_DBOW_SHIM_HPP
#define ORACLE_DBOW_SHIM_HPP
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "minicv.hpp"

namespace cv {
class FileNode {
public:
    FileNode operator[](const std::string &) const { std::abort(); }
    FileNode operator[](const char *) const { std::abort(); }
    FileNode operator[](int) const { std::abort(); }
    operator int() const { std::abort(); }
    operator double() const { std::abort(); }
    operator std::string() const { std::abort(); }
    size_t size() const { std::abort(); }
    bool empty() const { std::abort(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char *, int) { std::abort(); }
    FileStorage(const std::string &, int) { std::abort(); }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const std::string &) const { std::abort(); }
    FileNode operator[](const char *) const { std::abort(); }
};
template <typename T> static inline FileStorage &operator<<(FileStorage &fs, const T &) { std::abort(); return fs; }
}
#endif
