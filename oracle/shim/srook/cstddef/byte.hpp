// stand-in (oracle/shim/README.md): srook::byte == std::byte
#pragma once
#include <cstddef>
#include <srook/config/feature/constexpr.hpp>
namespace srook {
using byte = std::byte;
template <class I>
constexpr I to_integer(byte b) noexcept { return std::to_integer<I>(b); }
inline namespace literals {
inline namespace byte_literals {
constexpr byte operator"" _byte(unsigned long long v) noexcept { return byte(static_cast<unsigned char>(v)); }
}  // namespace byte_literals
}  // namespace literals
}  // namespace srook
