// stand-in (oracle/shim/README.md): srook::optional == std::optional (the payload policy is ignored)
#pragma once
#include <optional>
#include <srook/config/feature/constexpr.hpp>
namespace srook {
namespace optionally {
struct safe_optional_payload {};
}  // namespace optionally
template <class T, class Payload = optionally::safe_optional_payload>
struct optional : std::optional<T> {
    using std::optional<T>::optional;
    using std::optional<T>::operator=;
};
using std::nullopt;
using std::nullopt_t;
template <class T>
constexpr optional<std::decay_t<T>> make_optional(T&& v) { return optional<std::decay_t<T>>(std::forward<T>(v)); }
}  // namespace srook
