// stand-in (oracle/shim/README.md): srook::array == std::array
#pragma once
#include <array>
#include <srook/config/feature/constexpr.hpp>
namespace srook {
template <class T, std::size_t N>
using array = std::array<T, N>;
template <class T, class... Ts>
constexpr std::array<std::decay_t<T>, 1 + sizeof...(Ts)> make_array(T&& t, Ts&&... ts)
{
    return {{std::forward<T>(t), std::forward<Ts>(ts)...}};
}
}  // namespace srook
