// stand-in (oracle/shim/README.md) for srook::io::jpeg::bifstream -- decision O6 of SURVEY.md 8c.
// Whole-file big-endian reader.  Call sites: src/decoder/jpezy_decoder.hpp:68,195-276,286-357,370-461,489-491,589,612,634.
//   (s | Byte) >> x          one byte into an integer / byte / enum object
//   (s | Word) >> x          two bytes, high first
//   (s | Byte_n(n)) >> str   n bytes into a std::string
//   (s | Bytes) >> str       str.size() bytes into a std::string
//   (s | Bits(n)) >> i       n bits, most significant first; the 0x00 after a 0xFF data byte is dropped; i < 0 at the end
// Byte-oriented reads discard the unread bits of the current byte.  next_address() points at the next unread byte.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
#include <srook/cstddef/byte.hpp>

namespace srook {
namespace io {
namespace jpeg {

struct bifstream {
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

    explicit bifstream(const char* file)
    {
        if (std::FILE* fp = file ? std::fopen(file, "rb") : nullptr) {
            std::uint8_t tmp[1 << 16];
            for (std::size_t n; (n = std::fread(tmp, 1, sizeof tmp, fp)) > 0;) data_.insert(data_.end(), tmp, tmp + n);
            std::fclose(fp);
            ok_ = true;
        }
    }
    bifstream(const std::uint8_t* p, std::size_t n) : data_(p, p + n), ok_(true) {}
    explicit operator bool() const noexcept { return ok_; }
    const srook::byte* next_address() const noexcept { return reinterpret_cast<const srook::byte*>(data_.data()) + pos_; }
    void skip_byte(long long n)
    {
        left_ = 0;
        const long long p = static_cast<long long>(pos_) + n;
        pos_ = p < 0 ? 0 : (static_cast<std::size_t>(p) > data_.size() ? data_.size() : static_cast<std::size_t>(p));
    }

    unsigned get_byte()
    {
        left_ = 0;
        if (pos_ >= data_.size()) throw std::runtime_error("bifstream: end of data");
        return data_[pos_++];
    }
    int get_bit()
    {
        if (!left_) {
            if (pos_ >= data_.size()) return -1;
            cur_ = data_[pos_++];
            if (cur_ == 0xffu && pos_ < data_.size() && data_[pos_] == 0x00u) ++pos_;
            left_ = 8;
        }
        --left_;
        return int((cur_ >> left_) & 1u);
    }
    int get_bits(std::size_t n)
    {
        int v = 0;
        for (std::size_t i = 0; i < n; ++i) {
            const int b = get_bit();
            if (b < 0) return -1;
            v = (v << 1) | b;
        }
        return v;
    }

    template <class T>
    static void assign(T& dst, unsigned long long v)
    {
        if constexpr (std::is_enum_v<T>) dst = static_cast<T>(static_cast<std::underlying_type_t<T>>(v));
        else dst = static_cast<T>(v);
    }

    template <class Tag>
    struct proxy {
        bifstream& s;
        Tag tag;
        template <class T>
        proxy& operator>>(T& x)
        {
            if constexpr (std::is_same_v<Tag, Byte_tag>) {
                assign(x, s.get_byte());
            } else if constexpr (std::is_same_v<Tag, Word_tag>) {
                const unsigned hi = s.get_byte(), lo = s.get_byte();
                assign(x, (hi << 8) | lo);
            } else if constexpr (std::is_same_v<Tag, Byte_n>) {
                x.resize(tag.n);
                for (std::size_t i = 0; i < tag.n; ++i) x[i] = static_cast<char>(s.get_byte());
            } else if constexpr (std::is_same_v<Tag, Bytes_tag>) {
                for (auto& c : x) c = static_cast<char>(s.get_byte());
            } else {
                x = static_cast<T>(s.get_bits(tag.n));
            }
            return *this;
        }
    };
    template <class Tag>
    friend proxy<Tag> operator|(bifstream& s, const Tag& t) { return proxy<Tag>{s, t}; }

private:
    std::vector<std::uint8_t> data_;
    std::size_t pos_ = 0;
    unsigned cur_ = 0;
    int left_ = 0;
    bool ok_ = false;
};

}  // namespace jpeg
}  // namespace io
}  // namespace srook
