// stand-in (oracle/shim/README.md) for srook::io::jpeg::bofstream -- decisions O4, O5, O7 of SURVEY.md 8c.
// In-memory big-endian writer.  Call sites: src/encoder/jpezy_writer.hpp:15-110, src/encoder/jpezy_encoder.hpp:76,189-220.
//   (s | Byte) << v << w     one byte per value            (markers, lengths: no stuffing)
//   (s | Word) << v          two bytes, high first
//   (s | Byte_n(n)) << ptr   n bytes from a char pointer
//   (s | Bytes) << range     every element of a byte range
//   (s | Bits(n)) << v       the low n bits of v, most significant first; every completed byte equal to 0xFF is followed
//                            by 0x00 (T.81 B.1.1.5)
// A byte-oriented write that follows a partial byte first completes that byte with 1-bits, through the same stuffing path
// (T.81 F.1.2.3; the alternative -- zero fill -- is selectable with JPEZY_SHIM_PAD_ZERO for the comparison tests).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <ios>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace srook {
namespace io {
namespace jpeg {

struct bofstream {
    struct Byte_tag {};
    struct Word_tag {};
    struct Bytes_tag {};
    struct Byte_n {
        explicit constexpr Byte_n(std::size_t n_) : n(n_) {}
        std::size_t n;
    };
    struct Bits {
        explicit constexpr Bits(std::size_t n_) : n(n_) {}
        std::size_t n;
    };
    static constexpr Byte_tag Byte{};
    static constexpr Word_tag Word{};
    static constexpr Bytes_tag Bytes{};

    bofstream(std::size_t buffer_size, const char* file, std::ios::openmode = std::ios::out) : cap_(buffer_size), name_(file ? file : "")
    {
        fp_ = name_.empty() ? nullptr : std::fopen(name_.c_str(), "wb");
        buf_.reserve(buffer_size < (std::size_t(1) << 22) ? buffer_size : (std::size_t(1) << 22));
    }
    bofstream(const bofstream&) = delete;
    ~bofstream()
    {
        if (fp_) std::fclose(fp_);
    }
    explicit operator bool() const noexcept { return fp_ != nullptr; }
    std::size_t wrote_size() const noexcept { return buf_.size(); }
    void output_file()
    {
        if (!fp_) throw std::runtime_error("bofstream: file is not open");
        if (!buf_.empty() && std::fwrite(buf_.data(), 1, buf_.size(), fp_) != buf_.size()) throw std::runtime_error("bofstream: write failed");
        std::fflush(fp_);
    }
    const std::vector<std::uint8_t>& buffer() const noexcept { return buf_; }

    // ---- primitive operations ----
    void put_raw(unsigned v)
    {
        if (buf_.size() >= cap_) throw std::runtime_error("bofstream: buffer overflow");
        buf_.push_back(static_cast<std::uint8_t>(v));
    }
    void put_bits(std::size_t n, unsigned long long v)
    {
        for (std::size_t i = n; i-- > 0;) {
            acc_ = (acc_ << 1) | unsigned((v >> i) & 1u);
            if (++nacc_ == 8) {
                put_raw(acc_ & 0xffu);
                if ((acc_ & 0xffu) == 0xffu) put_raw(0);
                acc_ = 0, nacc_ = 0;
            }
        }
    }
    void align()
    {
#ifdef JPEZY_SHIM_PAD_ZERO
        if (nacc_) put_bits(8 - nacc_, 0);
#else
        if (nacc_) put_bits(8 - nacc_, 0xff);
#endif
    }
    void put_byte(unsigned v) { align(), put_raw(v & 0xffu); }

    template <class T>
    static unsigned long long as_integer(const T& v)
    {
        if constexpr (std::is_enum_v<T>) return static_cast<unsigned long long>(static_cast<std::underlying_type_t<T>>(v));
        else return static_cast<unsigned long long>(v);
    }

    template <class Tag>
    struct proxy {
        bofstream& s;
        Tag tag;
        template <class T>
        proxy& operator<<(const T& v)
        {
            if constexpr (std::is_same_v<Tag, Byte_tag>) {
                s.put_byte(unsigned(as_integer(v)));
            } else if constexpr (std::is_same_v<Tag, Word_tag>) {
                const unsigned long long w = as_integer(v);
                s.put_byte(unsigned(w >> 8)), s.put_byte(unsigned(w));
            } else if constexpr (std::is_same_v<Tag, Byte_n>) {
                const char* p = v;
                for (std::size_t i = 0; i < tag.n; ++i) s.put_byte(static_cast<unsigned char>(p[i]));
            } else if constexpr (std::is_same_v<Tag, Bytes_tag>) {
                for (const auto& e : v) s.put_byte(unsigned(as_integer(e)));
            } else {
                s.put_bits(tag.n, as_integer(v));
            }
            return *this;
        }
    };
    template <class Tag>
    friend proxy<Tag> operator|(bofstream& s, const Tag& t) { return proxy<Tag>{s, t}; }

private:
    std::size_t cap_;
    std::string name_;
    std::FILE* fp_ = nullptr;
    std::vector<std::uint8_t> buf_;
    unsigned acc_ = 0;
    unsigned nacc_ = 0;
};

}  // namespace jpeg
}  // namespace io
}  // namespace srook
