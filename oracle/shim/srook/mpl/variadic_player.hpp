// stand-in (oracle/shim/README.md)
#pragma once
namespace srook {
template <class...>
struct pack {};
}  // namespace srook
