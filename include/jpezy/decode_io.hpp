// include/jpezy/decode_io.hpp -- drop-in for jpezy::decode_io<Range> (src/decoder/decode_io.hpp:27-59): ASCII P3 writer for
// the first width*height samples of three planes, header "P3\n# Decoded by jpezy\nW H\n255\n".
#ifndef JPEZY_B200_DECODE_IO_HPP
#define JPEZY_B200_DECODE_IO_HPP

#include <algorithm>
#include <ostream>
#include <string>
#include <vector>

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
        const std::string head = "P3\n# Decoded by jpezy\n" + std::to_string(io.width) + " " + std::to_string(io.height) + "\n" + std::to_string(io.max_color) + "\n";
        ofs.write(head.data(), std::streamsize(head.size()));
        const std::size_t m = std::min(n, io.r_.size());
        // the ASCII formatting is the decoder CLI's bottleneck (12 bytes out per pixel): ranges of pixels on the host's cores
        const unsigned parts = pnm_detail::host_threads(m * 12);
        std::vector<std::string> chunk(parts);
        pnm_detail::parallel_parts(parts, [&](unsigned k) {
            pnm_detail::format_triples(io.r_, io.g_, io.b_, m * k / parts, m * (k + 1) / parts, chunk[k]);
        });
        for (const std::string& c : chunk) ofs.write(c.data(), std::streamsize(c.size()));
        return ofs;
    }
    const Range &r_, &g_, &b_;
};

template <class Range>
decode_io(std::size_t, std::size_t, const Range&, const Range&, const Range&) -> decode_io<Range>;

}  // namespace jpezy
#endif
