// stand-in (oracle/shim/README.md)
#pragma once
#include <type_traits>
namespace srook {
using std::is_arithmetic;
using std::is_same;
using std::remove_reference;
using std::remove_reference_t;
template <class T>
struct type_constant {
    using type = T;
};
template <class...>
using void_t = void;
namespace type_traits {
namespace detail {
template <class... B>
using Lor = std::disjunction<B...>;
template <class... B>
using Land = std::conjunction<B...>;
}  // namespace detail
}  // namespace type_traits
}  // namespace srook
