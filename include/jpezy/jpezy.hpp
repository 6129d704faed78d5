// include/jpezy/jpezy.hpp -- host-side common types of the drop-in: what src/jpezy.hpp of the reference gives its users
// (mode tags, MARKER, property + make_property, raii_messenger, disp_logo), restated over the standard library.
// The codec tables (ZZ, YQuantumTb, CQuantumTb, src/jpezy.hpp:36-45,131-152) are only needed by the header writer and
// live in jpezy_writer.hpp; the hot path has its own copies on the device (jpezy_b200/csrc/tables.h).
#ifndef JPEZY_B200_JPEZY_HPP
#define JPEZY_B200_JPEZY_HPP

#include <chrono>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <optional>
#include <string>
#include <utility>

namespace jpezy {

using byte = std::uint8_t;   // srook::byte in the reference: an opaque 8-bit type

// src/jpezy.hpp:20-29
inline void disp_logo()
{
    std::cout << "   _\n"
              << "  (_)_ __   ___ _____   _\n"
              << "  | | '_ \\ / _ \\_  / | | | \n"
              << "  | | |_) |  __// /| |_| |\n"
              << " _/ | .__/ \\___/___|\\__, |\n"
              << "|__/|_|             |___/\tby roki\n"
              << std::endl;
}

// src/jpezy.hpp:31-34
struct Release;
struct Debug;
struct COLOR_MODE;
struct GRAY_MODE;

// src/jpezy.hpp:47-127 (the values the host-side marker code needs)
enum class MARKER : std::uint8_t {
    SOF0 = 0xc0, SOF1 = 0xc1, SOF2 = 0xc2, SOF3 = 0xc3, DHT = 0xc4, SOF5 = 0xc5, SOF6 = 0xc6, SOF7 = 0xc7, JPG = 0xc8,
    SOF9 = 0xc9, SOF10 = 0xca, SOF11 = 0xcb, DAC = 0xcc, SOF13 = 0xcd, SOF14 = 0xce, SOF15 = 0xcf,
    RST0 = 0xd0, RST7 = 0xd7, SOI = 0xd8, EOI = 0xd9, SOS = 0xda, DQT = 0xdb, DNL = 0xdc, DRI = 0xdd, DHP = 0xde, EXP = 0xdf,
    APP0 = 0xe0, APP15 = 0xef, JPG0 = 0xf0, JPG13 = 0xfd, COM = 0xfe, TEM = 0x01, RESst = 0x02, RESnd = 0xbf, Marker = 0xff
};

// src/jpezy.hpp:154-342
struct property {
    enum class Format { undefined, JFIF, JFXX };
    enum class Units { undefined, dots_inch, dots_cm };
    enum class ExtensionCodes { undefined = 0, JPEG = 0x10, oneByte_pixel = 0x11, threeByte_pixel = 0x13 };
    enum AnalyzedResult { Yet = 0, is_htable = 0x01, is_qtable = 0x02, is_jfif = 0x04, is_comment = 0x08, is_start_data = 0x10 };
    enum class At {
        HSize, VSize, Dimension, SamplePrecision, Comment, Format, MajorRevisions, MinorRevisions, Units, HDensity, VDensity,
        HThumbnail, VThumbnail, ExtensionCode, Decodable, ELEMENT_SIZE
    };

    std::size_t width = 0, height = 0;
    int dimension = 0, sample_precision = 0;
    std::string comment;
    Format format = Format::undefined;
    byte major_rev = 0, minor_rev = 0;
    Units uni = Units::undefined;
    int width_density = 1, height_density = 1, width_thumbnail = 0, height_thumbnail = 0;
    ExtensionCodes ext = ExtensionCodes::undefined;
    int decodable = Yet;

    property() = default;
    explicit property(std::size_t w, std::size_t h, int dim, int sample_pre, std::string com, Format form, byte marev, byte mirev, Units u,
                      int wd, int hd, int wt, int ht, ExtensionCodes e, int decflag = AnalyzedResult::Yet)
        : width(w), height(h), dimension(dim), sample_precision(sample_pre), comment(std::move(com)), format(form), major_rev(marev),
          minor_rev(mirev), uni(u), width_density(wd), height_density(hd), width_thumbnail(wt), height_thumbnail(ht), ext(e),
          decodable(decflag)
    {
    }

    template <At at>
    const auto& get() const noexcept
    {
        return const_cast<property*>(this)->get<at>();
    }
    template <At at>
    auto& get() noexcept
    {
        if constexpr (at == At::HSize) return width;
        else if constexpr (at == At::VSize) return height;
        else if constexpr (at == At::Dimension) return dimension;
        else if constexpr (at == At::SamplePrecision) return sample_precision;
        else if constexpr (at == At::Comment) return comment;
        else if constexpr (at == At::Format) return format;
        else if constexpr (at == At::MajorRevisions) return major_rev;
        else if constexpr (at == At::MinorRevisions) return minor_rev;
        else if constexpr (at == At::Units) return uni;
        else if constexpr (at == At::HDensity) return width_density;
        else if constexpr (at == At::VDensity) return height_density;
        else if constexpr (at == At::HThumbnail) return width_thumbnail;
        else if constexpr (at == At::VThumbnail) return height_thumbnail;
        else if constexpr (at == At::ExtensionCode) return ext;
        else return decodable;
    }
};

// The reference builds `property` from Boost.Parameter named arguments (src/jpezy.hpp:346-386); the drop-in offers the
// same fifteen names as a plain aggregate with designated initialisers:
//   make_property({.width = w, .height = h, .dimension = 3, ...})
// (detail::make_property_impl swaps the two thumbnail arguments, src/jpezy.hpp:372 -- both are 0 at every call site)
struct property_args {
    std::size_t width = 0, height = 0;
    int dimension = 0, sample_precision = 0;
    std::string comment;
    property::Format format = property::Format::undefined;
    byte major_rev = 0, minor_rev = 0;
    property::Units units = property::Units::undefined;
    int width_density = 1, height_density = 1, width_thumbnail = 0, height_thumbnail = 0;
    property::ExtensionCodes extension_code = property::ExtensionCodes::undefined;
    int decodable = property::AnalyzedResult::Yet;
};
inline property make_property(const property_args& a)
{
    return property{a.width, a.height, a.dimension, a.sample_precision, a.comment, a.format, a.major_rev, a.minor_rev, a.units,
                    a.width_density, a.height_density, a.height_thumbnail, a.width_thumbnail, a.extension_code, a.decodable};
}

// src/jpezy.hpp:388-432: "<msg> " ... "Done! Processing time: X(sec)" with millisecond resolution of system_clock
struct raii_messenger {
    raii_messenger(const char* message, const char* ind = "") : mes(message), indent(ind), stoped(false)
    {
        std::cout << indent << mes << " ";
        start = std::chrono::system_clock::now();
    }
    void restart(const char* str = nullptr)
    {
        if (stoped) {
            if (str) std::cout << str << std::endl;
            else std::cout << mes << " ";
            start = std::chrono::system_clock::now();
            stoped = false;
        }
    }
    std::optional<float> stop()
    {
        if (!stoped) {
            end = std::chrono::system_clock::now();
            const float time = static_cast<float>(std::chrono::duration_cast<std::chrono::milliseconds>(end - start).count()) / 1000;
            std::cout << indent << "Done! Processing time: " << time << "(sec)" << std::endl;
            stoped = true;
            return time;
        }
        return std::nullopt;
    }
    ~raii_messenger() { stop(); }

private:
    std::chrono::system_clock::time_point start, end;
    const char *mes, *indent;
    bool stoped;
};

}  // namespace jpezy
#endif
