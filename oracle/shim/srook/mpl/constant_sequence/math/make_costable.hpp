// stand-in (oracle/shim/README.md), decision O1: value[u*B + x] = cos((2x+1) u pi / (2B)) in double.
// Layout: the only one consistent with DCT (src/encoder/jpezy_encoder.hpp:160) and inverse_dct
// (src/decoder/jpezy_decoder.hpp:664).  Evaluated by GCC's constant folder (correctly rounded).
#pragma once
#include <array>
#include <cstddef>
namespace srook {
namespace constant_sequence {
namespace math {
template <std::size_t A, std::size_t B>
struct make_costable_t {
    static constexpr std::size_t rows = A, cols = B;
};
namespace unwrap_costable {
template <class Table>
struct array;
template <std::size_t A, std::size_t B>
struct array<make_costable_t<A, B>> {
    static constexpr std::array<const double, A * B> make()
    {
        return make_impl(std::make_index_sequence<A * B>());
    }
    template <std::size_t... I>
    static constexpr std::array<const double, A * B> make_impl(std::index_sequence<I...>)
    {
        return {{__builtin_cos(double((2 * (I % B) + 1) * (I / B)) * 3.14159265358979323846 / double(2 * B))...}};
    }
    static constexpr std::array<const double, A * B> value = make();
};
}  // namespace unwrap_costable
}  // namespace math
}  // namespace constant_sequence
}  // namespace srook
