// jpezy_decode -- drop-in for the reference's decoder CLI (src/decoder/main.cpp:91-150):
//   jpezy_decode <input.(jpg | jpeg)> <output.ppm> [--gray] [-v]
// Same argv sniffing, console lines and exit codes; entropy decoding, IDCT and colour conversion run on the B200 through
// libjpezy_b200.so.  argv[3] / argv[4] are only looked at when they exist (the reference reads argv[argc], :99-104).
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <optional>
#include <stdexcept>
#include <string_view>

#include "jpezy/decode_io.hpp"
#include "jpezy/jpezy_decoder.hpp"

namespace {

int disp_error()
{
    std::cerr << "Usage: jpezy_decode <input.(jpg | jpeg)> ( <output.ppm | [OPT: --gray]> | -v )" << std::endl;
    return EXIT_FAILURE;
}

template <class CL, class T>
int output(jpezy::decoder<T>& dec, const char* out)
{
    auto raw_op = dec.template decode<CL>();
    if (!raw_op) {
        std::cerr << "decode failed" << std::endl;
        return EXIT_FAILURE;
    }
    const auto& [r, g, b] = raw_op.value();
    jpezy::decode_io dec_io(dec.pr.template get<jpezy::property::At::HSize>(), dec.pr.template get<jpezy::property::At::VSize>(), r, g, b);
    std::ofstream ofs(out, std::ios_base::out | std::ios_base::trunc);
    ofs << dec_io;
    std::cout << "Decoded image: "
              << "Netpbm image data, size = " << dec.pr.template get<jpezy::property::At::HSize>() << " x "
              << dec.pr.template get<jpezy::property::At::VSize>() << ", pixmap, ASCII text" << std::endl;
    return EXIT_SUCCESS;
}

template <class T>
int run(const char* in, const char* out, bool gray)
{
    jpezy::decoder<T> dec(in);
    return gray ? output<jpezy::GRAY_MODE>(dec, out) : output<jpezy::COLOR_MODE>(dec, out);
}

}  // namespace

int main(const int argc, const char* argv[])
{
    if (argc > 5 || argc < 3) return disp_error();
    const std::string_view sv0 = argv[1], sv1 = argv[2];
    std::optional<std::string_view> sv2, sv3;
    if (argc > 3) sv2 = argv[3];
    if (argc > 4) sv3 = argv[4];
    const auto has = [](const std::optional<std::string_view>& s, const char* what) { return s && s->find(what) != std::string_view::npos; };

    if (!((sv0.find("jpeg", sv0.find_first_of('.')) != std::string_view::npos || sv0.find("jpg", sv0.find_first_of('.')) != std::string_view::npos) &&
          sv1.find("ppm", sv1.find_first_of('.')) != std::string_view::npos))
        return disp_error();
    const bool gray = has(sv2, "--gray") || has(sv3, "--gray");
    const bool verbose = has(sv2, "-v") || has(sv3, "-v");   // note: "--gray" does not contain "-v"

    jpezy::disp_logo();
    try {
        return verbose ? run<jpezy::Debug>(argv[1], argv[2], gray) : run<jpezy::Release>(argv[1], argv[2], gray);
    } catch (const std::runtime_error& e) {   // no CUDA device: the reference has no such failure, report it like a failed decode
        std::cerr << e.what() << std::endl << "decode failed" << std::endl;
        return EXIT_FAILURE;
    }
}
