#pragma once
// oracle shim: plain (non-gzip) pass-through of an ifstream/ofstream.
#include <fstream>
#include <istream>
#include <ostream>
#include <stdexcept>
namespace boost { namespace iostreams {
struct gzip_decompressor {};
struct gzip_compressor {};
struct gzip_error : std::runtime_error { gzip_error() : std::runtime_error("gzip unsupported in oracle shim") {} };
struct filtering_istream : std::istream {
    filtering_istream() : std::istream(nullptr) {}
    void push(const gzip_decompressor&) { throw gzip_error(); }
    void push(std::istream& f) { rdbuf(f.rdbuf()); }
};
struct filtering_ostream : std::ostream {
    filtering_ostream() : std::ostream(nullptr) {}
    void push(const gzip_compressor&) { throw gzip_error(); }
    void push(std::ostream& f) { rdbuf(f.rdbuf()); }
};
typedef filtering_istream filtering_istream_t;
} }
