// Stand-in for boost::iostreams::filtering_istream + gzip_decompressor (Boost is not available in the build container).
// TEST INFRASTRUCTURE ONLY: lets the reference's own loadGridData (external/Alm/Alm_cpp/bilinear_interpol.cpp:27-99) read
// its gzip'ed grid files through zlib.  in.push(gzip_decompressor()); in.push(file) -> the inflated text becomes the stream.
#pragma once
#include <zlib.h>
#include <fstream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost { namespace iostreams {
struct gzip_decompressor {};
struct gzip_compressor {};
struct gzip_error : std::runtime_error { using std::runtime_error::runtime_error; };
class filtering_istream : public std::istringstream {
  public:
    void push(const gzip_decompressor&) {}
    void push(std::istream& src)
    {
        const std::string raw((std::istreambuf_iterator<char>(src)), std::istreambuf_iterator<char>());
        z_stream zs{};
        if (inflateInit2(&zs, 15 + 32) != Z_OK) throw gzip_error("inflateInit2");
        zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(raw.data()));
        zs.avail_in = (uInt)raw.size();
        std::string out;
        char buf[1 << 15];
        int rc = Z_OK;
        while (rc == Z_OK) {
            zs.next_out = reinterpret_cast<Bytef*>(buf); zs.avail_out = sizeof(buf);
            rc = inflate(&zs, Z_NO_FLUSH);
            out.append(buf, sizeof(buf) - zs.avail_out);
        }
        inflateEnd(&zs);
        if (rc != Z_STREAM_END) throw gzip_error("inflate");
        this->str(out);
    }
};
}}  // namespace boost::iostreams
