// include/jpezy/jpezy_decoder.hpp -- drop-in for jpezy::decoder<BuildMode> (src/decoder/jpezy_decoder.hpp:39-136).
// Same constructor (file name), same decode<MODE_TAG>() -> optional<array<vector<byte>,3>>, public `pr`, same console
// lines (Debug = the reference's verbose -v trace).  Marker parsing (analyze_*, get_marker, :171-502) is restated here on
// the host; the MCU loop (:107-130: decode_huffman, inverse_quantization, inverse_dct, decode_mcu, make_rgb) runs on the
// B200 behind jpezyb200_decode.
#ifndef JPEZY_B200_JPEZY_DECODER_HPP
#define JPEZY_B200_JPEZY_DECODER_HPP

#include <array>
#include <cstdio>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "jpezy.hpp"
#include "runtime.hpp"

namespace jpezy {

template <class BuildMode = Release>
struct decoder {
    explicit decoder(const char* filename)
    {
        pr.width_density = pr.height_density = 1;   // src/decoder/jpezy_decoder.hpp:42-59
        std::memset(&frame, 0, sizeof frame);
        if (std::FILE* fp = filename ? std::fopen(filename, "rb") : nullptr) {
            byte tmp[1 << 16];
            for (std::size_t n; (n = std::fread(tmp, 1, sizeof tmp, fp)) > 0;) file.insert(file.end(), tmp, tmp + n);
            std::fclose(fp);
        }
    }

    static constexpr bool is_release_mode = std::is_same_v<Release, BuildMode>;
    static constexpr std::size_t rgb_size = 3, block_size = 8, blocks_size = 64, mcu_size = 4;

    template <class MODE_TAG = COLOR_MODE>
    std::optional<std::array<std::vector<byte>, 3>> decode()
    {
        raii_messenger mes("process started...");
        std::cout << '\n';
        try {
            analyze_header();
        } catch (const std::runtime_error&) {
            return {};
        }
        disp_info("\t");
        if (!(pr.decodable & (property::is_htable | property::is_qtable | property::is_start_data))) return {};
        std::unique_ptr<raii_messenger> mes_dec;
        if constexpr (!is_release_mode) mes_dec = std::make_unique<raii_messenger>("decoding started...", "\t");

        frame.width = static_cast<std::uint32_t>(pr.width), frame.height = static_cast<std::uint32_t>(pr.height);
        frame.sample_precision = static_cast<std::uint8_t>(pr.sample_precision), frame.ncomp = static_cast<std::uint8_t>(pr.dimension);
        frame.restart_interval = static_cast<std::uint16_t>(restart_interval);
        const std::size_t plane = jpezyb200_plane_bytes(&frame);   // src/decoder/jpezy_decoder.hpp:94-101
        std::array<std::vector<byte>, 3> rgb;
        for (auto& v : rgb) v.resize(plane);
        if (pos >= file.size()) {
            std::cerr << "decode_mcu(): throw exception from decode_huffman" << std::endl;
            return {};
        }
        const int rc = jpezyb200_decode(b200::runtime::ctx(), file.data() + pos, file.size() - pos, &frame, std::is_same_v<MODE_TAG, GRAY_MODE> ? 1 : 0,
                                        rgb[0].data(), rgb[1].data(), rgb[2].data(), plane);
        if (rc != JPEZYB200_OK) {
            // the reference reports a failing MCU like this and returns an empty optional (:109-114)
            std::cerr << "decode_mcu(): throw exception from " << jpezyb200_strerror(rc) << " (" << jpezyb200_last_error(b200::runtime::ctx()) << ")" << std::endl;
            return {};
        }
        return {std::move(rgb)};
    }

    property pr;

private:
    // ---- byte reader over the whole file (the marker-level part of srook::io::jpeg::bifstream) ----
    unsigned get_byte()
    {
        if (pos >= file.size()) throw std::runtime_error("bifstream: end of data");
        return file[pos++];
    }
    unsigned get_word()
    {
        const unsigned hi = get_byte();
        return (hi << 8) | get_byte();
    }
    void skip_byte(long long n)
    {
        const long long p = static_cast<long long>(pos) + n;
        pos = p < 0 ? 0 : (static_cast<std::size_t>(p) > file.size() ? file.size() : static_cast<std::size_t>(p));
    }

    // src/decoder/jpezy_decoder.hpp:139-150
    void disp_info(const char* indent = "")
    {
        std::cout << indent << "Loaded JPEG: " << pr.width << "x" << pr.height << ", "
                  << "presicion " << pr.sample_precision << ", "
                  << "\"" << pr.comment << "\""
                  << ", " << (pr.format == property::Format::JFIF ? "JFIF" : pr.format == property::Format::JFXX ? "JFXX" : "undefined") << " standart "
                  << static_cast<unsigned>(pr.major_rev) << ".0" << static_cast<unsigned>(pr.minor_rev) << ", "
                  << (pr.uni == property::Units::dots_inch ? "dots inch" : pr.uni == property::Units::dots_cm ? "dots cm" : "undefined") << ", "
                  << "frames " << pr.dimension << ", "
                  << "density " << pr.width_density << "x" << pr.height_density << "\n"
                  << std::endl;
    }

    std::unique_ptr<raii_messenger> trace(const char* msg, const char* indent = "")
    {
        if constexpr (!is_release_mode) return std::make_unique<raii_messenger>(msg, indent);
        else return nullptr;
    }
    void found(const char* name, bool leading_newline = false)
    {
        if constexpr (!is_release_mode) std::cout << (leading_newline ? "\n" : "") << "\t\tfound marker: [" << name << "]" << std::endl;
    }

    // :171-188
    void analyze_header()
    {
        auto mes = trace("analyzing header...", "\t");
        do {
            if (get_marker() == MARKER::SOI) enable = true;
        } while (!enable);
        while (enable) {
            pr.decodable |= analyze_marker();
            if (pr.decodable & property::is_start_data) return;
        }
        throw std::runtime_error(__func__);
    }

    // :190-256 -- BITS / HUFFVAL are kept as they are (the device builds its own canonical decoder tables from them)
    void analyze_dht(std::size_t size)
    {
        auto mes = trace("analyzing DHT...", "\t\t\t");
        const std::size_t end = pos + size;
        do {
            const unsigned uc = get_byte();
            const unsigned tc = uc >> 4, th = uc & 0x0f;
            if (tc > 1) throw std::runtime_error("DC format error");
            if (th > 3) throw std::runtime_error("AC format error");
            jpezyb200_huff& h = frame.ht[tc][th];
            std::memset(&h, 0, sizeof h);
            std::size_t n = 0;
            for (int i = 0; i < 16; ++i) h.bits[i] = static_cast<std::uint8_t>(get_byte()), n += h.bits[i];
            if constexpr (!is_release_mode) std::cout << " size: " << n << std::endl;
            for (std::size_t k = 0; k < n; ++k) {
                const unsigned v = get_byte();
                if (k < 256) h.vals[k] = static_cast<std::uint8_t>(v);
            }
            h.present = 1;
            if constexpr (!is_release_mode) {
                switch (n) {
                case 12: std::cout << "found DC Huffman Table... "; break;
                case 162: std::cout << "found AC Huffman Table... "; break;
                default: std::cerr << n << std::endl; throw std::out_of_range("invalid size table");
                }
            }
        } while (pos < end);
    }

    // :258-277 -- tables are de-zig-zagged on load
    void analyze_dqt(std::size_t size)
    {
        auto mes = trace("\t\t\tanalyzing DQT...");
        const std::size_t end = pos + size;
        static constexpr int ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                       41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                       30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        do {
            const unsigned c = get_byte();
            std::uint16_t* q = frame.qt[c & 0x3];
            if (!(c >> 4)) {
                for (int i = 0; i < 64; ++i) q[ZZ[i]] = static_cast<std::uint16_t>(get_byte());
            } else {
                for (int i = 0; i < 64; ++i) q[ZZ[i]] = static_cast<std::uint16_t>(get_word());
            }
        } while (pos < end);
    }

    // :279-305
    void analyze_frame()
    {
        auto mes = trace("\t\t\tanalyzing frames...");
        pr.sample_precision = int(get_byte());
        pr.height = get_word();
        pr.width = get_word();
        pr.dimension = int(get_byte());
        if (pr.dimension != 3 && pr.dimension != 1) throw std::runtime_error("Sorry, this dimension size is not supported");
        if constexpr (!is_release_mode) std::cout << "VSize: " << pr.height << " HSize: " << pr.width << " ";
        for (int i = 0; i < pr.dimension; ++i) {
            get_byte();   // component identifier C
            const unsigned c = get_byte();
            frame.hs[i] = static_cast<std::uint8_t>(c >> 4), frame.vs[i] = static_cast<std::uint8_t>(c & 0xf);
            frame.tq[i] = static_cast<std::uint8_t>(get_byte());
        }
    }

    // :307-334
    void analyze_scan()
    {
        auto mes = trace("\t\t\tanalyzing scan data...");
        const unsigned ncomp = get_byte();
        for (unsigned i = 0; i < ncomp; ++i) {
            get_byte();   // Cs
            const unsigned c = get_byte();
            if (i < 3) frame.td[i] = static_cast<std::uint8_t>(c >> 4), frame.ta[i] = static_cast<std::uint8_t>(c & 0xf);
            if ((c >> 4) > 2 || (c & 0xf) > 2) throw std::out_of_range(__func__);
        }
        get_byte(), get_byte(), get_byte();   // Ss, Se, Ah/Al: unused for sequential DCT
    }

    // :336-357
    void analyze_jfif()
    {
        auto mes = trace("\t\t\tanalyzing jfif...");
        pr.format = property::Format::JFIF;
        pr.major_rev = static_cast<byte>(get_byte()), pr.minor_rev = static_cast<byte>(get_byte());
        pr.uni = static_cast<property::Units>(get_byte());
        pr.width_density = int(get_word()), pr.height_density = int(get_word());
        pr.width_thumbnail = int(get_byte()), pr.height_thumbnail = int(get_byte());
        pr.decodable |= property::is_jfif;
    }
    void analyze_jfxx()
    {
        auto mes = trace("\t\t\tanalyzing jfxx...");
        pr.format = property::Format::JFXX;
        pr.ext = static_cast<property::ExtensionCodes>(get_byte());
    }

    // :360-484
    int analyze_marker()
    {
        int length = 0;
        const unsigned m = static_cast<unsigned>(get_marker());
        switch (m) {
        case 0xc0: found("SOF0"), length = int(get_word()), analyze_frame(); break;
        case 0xc4: found("DHT"), length = int(get_word()) - 2, analyze_dht(std::size_t(length)); return property::is_htable;
        case 0xdc: found("DNL"), length = int(get_word()), pr.height = get_word(); break;
        case 0xdb: found("DQT"), length = int(get_word()) - 2, analyze_dqt(std::size_t(length)); return property::is_qtable;
        case 0xd9: found("EOI"), enable = false; break;
        case 0xda: found("SOS"), length = int(get_word()), analyze_scan(); return property::is_start_data;
        case 0xdd: found("DRI"), length = int(get_word()), restart_interval = get_word(); break;
        case 0xfe:
            found("COM"), length = int(get_word()) - 2;
            pr.comment.resize(std::size_t(length < 0 ? 0 : length));
            for (char& ch : pr.comment) ch = static_cast<char>(get_byte());
            return property::is_comment;
        case 0xc1: case 0xc2: case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce:
        case 0xcf: case 0xdf: case 0xcc: case 0xde:
            break;   // the reference constructs a runtime_error("Not supported") without throwing it (:419)
        case 0xe0: {
            found("APP0", true), length = int(get_word()) - 2;
            if (length >= 4) {
                std::string id(5, '\0');
                for (char& ch : id) ch = static_cast<char>(get_byte());
                id.pop_back();
                if (id == "JFIF") analyze_jfif(), skip_byte(length - 14);
                else if (id == "JFXX") analyze_jfxx(), skip_byte(length - 1);
                else skip_byte(length - 4);
            } else {
                skip_byte(length);
            }
            break;
        }
        default:
            if (m >= 0xe1 && m <= 0xef) {
                length = int(get_word()) - 2, skip_byte(length);
                break;
            }
            throw std::runtime_error("Marker error");
        }
        return property::Yet;
    }

    // :486-502
    MARKER get_marker()
    {
        for (;;) {
            unsigned c = get_byte();
            if (c == 0xff) {
                c = get_byte();
                if (c) {
                    if (c > 0x02 && c < 0xc0) return MARKER::Marker;   // MARKER::Error == 0xff
                    return static_cast<MARKER>(c);
                }
            }
        }
    }

    std::vector<byte> file;
    std::size_t pos = 0;
    std::size_t restart_interval = 0;
    jpezyb200_frame frame;
    bool enable = false;
};

}  // namespace jpezy
#endif
