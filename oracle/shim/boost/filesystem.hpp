// Minimal stand-in for <boost/filesystem.hpp> -- TEST INFRASTRUCTURE ONLY.
//
// Only what /root/reference/src/IO/FSPath.cpp and inc/IO/FSPath.hpp use.  boost's
// path::string() returns a const reference to the stored string (FSPath::str() forwards that
// reference), which std::filesystem::path does not -- hence this small class on top of it.
#pragma once
#include <cassert>
#include <filesystem>
#include <string>
namespace boost { namespace filesystem {
class path {
public:
    path() {}
    path(const std::string& s) : _s(s) {}
    path(const char* s) : _s(s) {}
    const std::string& string() const { return _s; }
    const char* c_str() const { return _s.c_str(); }
    path& operator/=(const path& other) {
        _s = (std::filesystem::path(_s) / std::filesystem::path(other._s)).string();
        return *this;
    }
    path parent_path() const { return path(std::filesystem::path(_s).parent_path().string()); }
private:
    std::string _s;
};
inline bool exists(const path& p) { return std::filesystem::exists(p.string()); }
inline bool create_directories(const path& p) {
    std::error_code ec;
    return std::filesystem::create_directories(p.string(), ec);
}
}}
