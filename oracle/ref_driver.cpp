// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN encoder / decoder classes,
// compiled unmodified from /root/reference/src against oracle/shim (see oracle/shim/README.md for what that pins and
// what it does not).  Output: oracle/_ref/libjpezy_ref.so (git-ignored).  Used by tests/golden/make_golden.py and
// tests/test_oracle_vs_ref.py to validate the restatement in jpezy_oracle.cpp; never by the product.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "encoder/jpezy_encoder.hpp"
#include "decoder/jpezy_decoder.hpp"

namespace {
struct quiet {      // the reference prints progress lines to std::cout / std::cerr
    std::ostringstream sink;      // declared first: members are initialised in declaration order
    std::streambuf *o, *e;
    quiet() : o(std::cout.rdbuf(sink.rdbuf())), e(std::cerr.rdbuf(sink.rdbuf())) {}
    ~quiet() { std::cout.rdbuf(o), std::cerr.rdbuf(e); }
};
}  // namespace

extern "C" {

// jpezy::encoder<srook::byte>(pr, r, g, b).encode<MODE>(path) exactly as encode_io's operator<< drives it
// (src/encoder/encode_io.hpp:144-166 colour, :177-193 gray).  Returns wrote_size(), -1 on exception.
long long ref_encode_file(const std::uint8_t* r, const std::uint8_t* g, const std::uint8_t* b, int W, int H, int gray, const char* path)
{
    quiet q;
    try {
        const std::size_t n = std::size_t(W) * std::size_t(H);
        std::vector<srook::byte> rv(n), gv(n), bv(n);
        for (std::size_t i = 0; i < n; ++i) rv[i] = srook::byte(r[i]), gv[i] = srook::byte(g[i]), bv[i] = srook::byte(b[i]);
        const jpezy::property pr(std::size_t(W), std::size_t(H), 3, 8, gray ? "Encoded by JPEZY" : "Encoded by jpezy", jpezy::property::Format::JFIF,
                                 srook::byte(1), srook::byte(2), jpezy::property::Units::dots_inch, 96, 96, 0, 0,
                                 jpezy::property::ExtensionCodes::undefined);
        jpezy::encoder enc(pr, rv, gv, bv);
        return static_cast<long long>(gray ? enc.encode<jpezy::GRAY_MODE>(path) : enc.encode<jpezy::COLOR_MODE>(path));
    } catch (const std::exception&) {
        return -1;
    }
}

// jpezy::decoder<Release>(path).decode<MODE>().  Returns 0 and fills W, H, plane_len; planes are copied when the
// pointers are non-null and cap >= plane_len.  1 = decode() returned an empty optional, 2 = capacity.
int ref_decode_file(const char* path, int gray, int* W, int* H, std::size_t* plane_len, std::uint8_t* r, std::uint8_t* g, std::uint8_t* b, std::size_t cap)
{
    quiet q;
    try {
        jpezy::decoder<jpezy::Release> dec(path);
        auto res = gray ? dec.decode<jpezy::GRAY_MODE>() : dec.decode<jpezy::COLOR_MODE>();
        if (!res) return 1;
        const auto& planes = res.value();
        if (W) *W = int(dec.pr.get<jpezy::property::At::HSize>());
        if (H) *H = int(dec.pr.get<jpezy::property::At::VSize>());
        if (plane_len) *plane_len = planes[0].size();
        if (r && g && b) {
            if (cap < planes[0].size()) return 2;
            std::memcpy(r, planes[0].data(), planes[0].size());
            std::memcpy(g, planes[1].data(), planes[1].size());
            std::memcpy(b, planes[2].data(), planes[2].size());
        }
        return 0;
    } catch (const std::exception&) {
        return 3;
    }
}

// the constants the third-party stand-ins produce (decisions O1, O2), for comparison with the oracle's
void ref_constants(double* cos_table64, double* dis_sqrt)
{
    constexpr auto t = srook::constant_sequence::math::unwrap_costable::array<srook::constant_sequence::math::make_costable_t<8, 8>>::value;
    for (int i = 0; i < 64; ++i) cos_table64[i] = t[i];
    *dis_sqrt = 1.0 / srook::sqrt(2.0);
}

}  // extern "C"

// Command-line face of the two entry points above, so that the reference's DECODER can run in a process of its own:
// analyze_dht (src/decoder/jpezy_decoder.hpp:231) reads sizeTP[n], one element past the end of the vector, before it tests
// k >= n, and continues (writing codeTP[k]) when that stray word happens to equal the current code length.  In a fresh
// process the word is zero and nothing happens -- which is how the reference's own CLI gets away with it -- but inside a
// long-lived Python process it sees recycled heap memory (AddressSanitizer: heap-buffer-overflow, READ of size 8).
//   ref_tool decode <in.jpg> <gray 0|1> <out.bin>     out.bin = int32 W, int32 H, uint64 plane_len, then the R, G, B planes
#ifdef JPEZY_REF_TOOL
int main(int argc, char** argv)
{
    if (argc == 5 && std::string(argv[1]) == "decode") {
        int W = 0, H = 0;
        std::size_t pl = 0;
        const int gray = std::atoi(argv[3]);
        int rc = ref_decode_file(argv[2], gray, &W, &H, &pl, nullptr, nullptr, nullptr, 0);
        if (rc) return 10 + rc;
        std::vector<std::uint8_t> r(pl), g(pl), b(pl);
        rc = ref_decode_file(argv[2], gray, &W, &H, &pl, r.data(), g.data(), b.data(), pl);
        if (rc) return 10 + rc;
        std::FILE* fp = std::fopen(argv[4], "wb");
        if (!fp) return 2;
        const std::int32_t wh[2] = {W, H};
        const std::uint64_t n = pl;
        std::fwrite(wh, sizeof wh, 1, fp), std::fwrite(&n, sizeof n, 1, fp);
        std::fwrite(r.data(), 1, pl, fp), std::fwrite(g.data(), 1, pl, fp), std::fwrite(b.data(), 1, pl, fp);
        std::fclose(fp);
        return 0;
    }
    // ref_tool bench <in.rgb> <W> <H> <gray 0|1> <reps> <scratch.jpg>: times encoder::encode() (header + MCU loop + EOI + the
    // single fwrite of the buffer, src/encoder/jpezy_encoder.hpp:38-77) and decoder::decode() (marker parse + MCU loop,
    // src/decoder/jpezy_decoder.hpp:76-134) the way the CLIs call them; PNM parsing / writing is not part of either.
    // in.rgb = the three planes back to back.  Prints "<encode seconds> <decode seconds>" (sums over reps).
    if (argc == 8 && std::string(argv[1]) == "bench") {
        const int W = std::atoi(argv[3]), H = std::atoi(argv[4]), gray = std::atoi(argv[5]), reps = std::atoi(argv[6]);
        const std::size_t n = std::size_t(W) * std::size_t(H);
        std::vector<std::uint8_t> in(3 * n);
        std::FILE* fp = std::fopen(argv[2], "rb");
        if (!fp || std::fread(in.data(), 1, in.size(), fp) != in.size()) return 2;
        std::fclose(fp);
        double te = 0, td = 0;
        for (int i = 0; i < reps; ++i) {
            const auto t0 = std::chrono::steady_clock::now();
            if (ref_encode_file(in.data(), in.data() + n, in.data() + 2 * n, W, H, gray, argv[7]) < 0) return 3;
            const auto t1 = std::chrono::steady_clock::now();
            int w2 = 0, h2 = 0;
            std::size_t pl = 0;
            if (ref_decode_file(argv[7], gray, &w2, &h2, &pl, nullptr, nullptr, nullptr, 0)) return 4;
            const auto t2 = std::chrono::steady_clock::now();
            te += std::chrono::duration<double>(t1 - t0).count(), td += std::chrono::duration<double>(t2 - t1).count();
        }
        std::printf("%.6f %.6f\n", te, td);
        return 0;
    }
    std::fprintf(stderr, "usage: ref_tool decode <in.jpg> <gray 0|1> <out.bin> | bench <in.rgb> <W> <H> <gray> <reps> <scratch.jpg>\n");
    return 2;
}
#endif
