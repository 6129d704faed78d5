// include/jpezy/decode_io.hpp -- drop-in for jpezy::decode_io<Range> (src/decoder/decode_io.hpp:27-59): ASCII P3 writer for
// the first width*height samples of three planes, header "P3\n# Decoded by jpezy\nW H\n255\n".
#ifndef JPEZY_B200_DECODE_IO_HPP
#define JPEZY_B200_DECODE_IO_HPP

#include <ostream>
#include <string>

#include "pnm_stream.hpp"

namespace jpezy {

template <class Range>
struct decode_io : pnm_stream {
    decode_io(std::size_t w, std::size_t h, const Range& r, const Range& g, const Range& b) : pnm_stream(true, w, h, 255), r_(r), g_(g), b_(b)
    {
        if (!(r.size() == g.size() && g.size() == b.size())) initializing_succeed = false;
    }

private:
    friend std::ostream& operator<<(std::ostream& ofs, const decode_io& io)
    {
        if (!io.initializing_succeed) io.report_error(__func__);
        const std::size_t n = io.width * io.height;
        std::string out = "P3\n# Decoded by jpezy\n" + std::to_string(io.width) + " " + std::to_string(io.height) + "\n" + std::to_string(io.max_color) + "\n";
        out.reserve(out.size() + n * 12);
        char tmp[16];
        const auto put = [&](unsigned v, char sep) {
            int k = 0;
            do tmp[k++] = char('0' + v % 10), v /= 10; while (v);
            while (k) out.push_back(tmp[--k]);
            out.push_back(sep);
        };
        for (std::size_t i = 0; i < n && i < io.r_.size(); ++i) put(unsigned(io.r_[i]), ' '), put(unsigned(io.g_[i]), ' '), put(unsigned(io.b_[i]), '\n');
        ofs.write(out.data(), std::streamsize(out.size()));
        return ofs;
    }
    const Range &r_, &g_, &b_;
};

template <class Range>
decode_io(std::size_t, std::size_t, const Range&, const Range&, const Range&) -> decode_io<Range>;

}  // namespace jpezy
#endif
