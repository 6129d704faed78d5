// include/jpezy/encode_io.hpp -- drop-in for jpezy::encode_io, to_jpeg, gray_scale (src/encoder/encode_io.hpp:33-209):
//   jpezy::encode_io pnm("in.ppm");  ofs << (pnm | jpezy::to_jpeg("out.jpg"));  ofs << (pnm | to_jpeg(f) | jpezy::gray_scale);
// The P3 grammar is the reference's, quirks included: lines containing '#' are dropped wherever
// they appear, the size line must split into exactly two tokens, maxval is parsed and ignored, a final line without a
// newline is dropped, an interior empty token makes std::stoi throw std::invalid_argument.  One pass over the file
// instead of std::list<std::string> per token (the reference spends 0.52 of 0.57 s here, README.md:48-56).
#ifndef JPEZY_B200_ENCODE_IO_HPP
#define JPEZY_B200_ENCODE_IO_HPP

#include <cctype>
#include <fstream>
#include <iostream>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "jpezy.hpp"
#include "jpezy_encoder.hpp"
#include "pnm_stream.hpp"

namespace jpezy {

struct to_jpeg {
    explicit constexpr to_jpeg(const char* file_) : file(file_) {}
    const char* file;
};
struct gray_scale_t {};
inline constexpr gray_scale_t gray_scale{};

struct encode_io : pnm_stream {
    encode_io(const char* file_name) : pnm_stream(true, 0, 0, 0)
    {
        std::ifstream ifs(file_name);
        if (!ifs) {
            initializing_succeed = false;
            return;
        }
        // src/encoder/encode_io.hpp:50-56: read lines until one without '#' (or until getline fails; the last string read is returned)
        const auto jump_comment = [&ifs]() {
            std::string str;
            while (std::getline(ifs, str) && str.find('#') != std::string::npos) {}
            return str;
        };
        // boost::split(is_space()) with token_compress_off: every separator ends a token
        const auto split = [](const std::string& s) {
            std::vector<std::string> out(1);
            for (char c : s) {
                if (std::isspace(static_cast<unsigned char>(c))) out.emplace_back();
                else out.back().push_back(c);
            }
            return out;
        };
        std::string format = jump_comment();
        if (format != "P3") {
            initializing_succeed = false;
            return;
        }
        format = jump_comment();
        const std::vector<std::string> wh = split(format);
        if (wh.size() != 2) {
            initializing_succeed = false;
        } else {
            width = std::size_t(std::stoi(wh[0])), height = std::size_t(std::stoi(wh[1]));
            format = jump_comment();
            max_color = std::size_t(std::stoi(format));
            std::vector<value_type> img;
            img.reserve(width * height * 3);
            for (std::string line = jump_comment(); !ifs.eof(); line = jump_comment()) {
                // tokens of the line; exactly one trailing empty token is forgiven (:83-84), every other one reaches stoi
                std::size_t i = 0;
                const std::size_t n = line.size();
                for (;;) {
                    std::size_t j = i;
                    while (j < n && !std::isspace(static_cast<unsigned char>(line[j]))) ++j;
                    const bool last = j >= n;
                    if (!(last && j == i)) img.push_back(value_type(std::stoi(line.substr(i, j - i))));   // "" -> std::invalid_argument
                    if (last) break;
                    i = j + 1;
                }
            }
            rgb_img.resize(img.size() / 3);
            for (std::size_t k = 0; k < rgb_img.size(); ++k) rgb_img[k] = {img[3 * k], img[3 * k + 1], img[3 * k + 2]};
        }
        std::cout << "width: " << width << " height: " << height << std::endl;
    }

private:
    // P3 echo (:104-119)
    friend std::ostream& operator<<(std::ostream& os, const encode_io& pnm)
    {
        pnm.report_error(__func__);
        os << "P3\n" << pnm.width << " " << pnm.height << "\n" << pnm.max_color << "\n";
        for (const auto& rgb : pnm.rgb_img) os << unsigned(rgb[0]) << " " << unsigned(rgb[1]) << " " << unsigned(rgb[2]) << '\n';
        return os;
    }

    // :121-133
    std::tuple<std::vector<rgb_type>, std::vector<rgb_type>, std::vector<rgb_type>> split_rgb() const
    {
        std::vector<rgb_type> r(rgb_img.size()), g(rgb_img.size()), b(rgb_img.size());
        for (std::size_t i = 0; i < rgb_img.size(); ++i) r[i] = rgb_img[i][0], g[i] = rgb_img[i][1], b[i] = rgb_img[i][2];
        return {std::move(r), std::move(g), std::move(b)};
    }

    // :135-169
    friend std::ofstream& operator<<(std::ofstream& ofs, const std::pair<const to_jpeg, const encode_io&>& pnm)
    {
        ofs.close();
        pnm.second.report_error(__func__);
        const property pr = make_property({.width = pnm.second.width, .height = pnm.second.height, .dimension = 3, .sample_precision = 8,
                                           .comment = "Encoded by jpezy", .format = property::Format::JFIF, .major_rev = 1, .minor_rev = 2,
                                           .units = property::Units::dots_inch, .width_density = 96, .height_density = 96,
                                           .width_thumbnail = 0, .height_thumbnail = 0,
                                           .extension_code = property::ExtensionCodes::undefined, .decodable = property::Yet});
        auto [r, g, b] = pnm.second.split_rgb();
        encoder enc(pr, r, g, b);
        const std::size_t size = enc.encode<COLOR_MODE>(pnm.first.file);
        std::cout << "Output size: " << size << " byte" << std::endl;
        return ofs;
    }
    // :171-196
    friend std::ofstream& operator<<(std::ofstream& ofs, const std::pair<gray_scale_t, std::pair<const to_jpeg, const encode_io&>>& pnm)
    {
        ofs.close();
        pnm.second.second.report_error(__func__);
        const property pr{pnm.second.second.width, pnm.second.second.height, 3, 8, "Encoded by JPEZY", property::Format::JFIF, byte(1), byte(2),
                          property::Units::dots_inch, 96, 96, 0, 0, property::ExtensionCodes::undefined};
        auto [r, g, b] = pnm.second.second.split_rgb();
        encoder enc(pr, r, g, b);
        const std::size_t size = enc.encode<GRAY_MODE>(pnm.second.first.file);
        std::cout << "Output size: " << size << " srook::byte" << std::endl;   // sic (:193)
        return ofs;
    }
    friend std::pair<gray_scale_t, std::pair<const to_jpeg, const encode_io&>> operator|(const std::pair<const to_jpeg, const encode_io&>& pnm, const gray_scale_t& gr)
    {
        return {gr, pnm};
    }
    friend std::pair<const to_jpeg, const encode_io&> operator|(const encode_io& pnm, const to_jpeg& jpeg_tag) noexcept
    {
        return {jpeg_tag, pnm};
    }
};

}  // namespace jpezy
#endif
