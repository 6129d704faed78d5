// stand-in (oracle/shim/README.md)
#pragma once
namespace srook {
constexpr double cos(double x) { return __builtin_cos(x); }
}  // namespace srook
