// include/jpezy/pnm_stream.hpp -- base of the two PNM adaptors (src/pnm_stream.hpp:11-43)
#ifndef JPEZY_B200_PNM_STREAM_HPP
#define JPEZY_B200_PNM_STREAM_HPP

#include <array>
#include <stdexcept>
#include <string>
#include <vector>

#include "jpezy.hpp"

namespace jpezy {

struct pnm_stream {
    pnm_stream() : initializing_succeed(true), width(0), height(0), max_color(0) {}
    pnm_stream(bool b, std::size_t w, std::size_t h, std::size_t max) : initializing_succeed(b), width(w), height(h), max_color(max) {}
    explicit operator bool() const noexcept { return initializing_succeed; }

protected:
    typedef byte value_type;
    typedef byte rgb_type;
    bool initializing_succeed;
    std::size_t width, height, max_color;
    std::vector<std::array<rgb_type, 3>> rgb_img;

    void report_error(const char* funcName) const
    {
        if (initializing_succeed) return;
        throw std::runtime_error(std::string("Initializing was failed: ") + funcName);
    }
};

}  // namespace jpezy
#endif
