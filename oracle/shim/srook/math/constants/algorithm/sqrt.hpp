// stand-in (oracle/shim/README.md), decision O2: IEEE square root, usable in constant expressions
#pragma once
namespace srook {
constexpr double sqrt(double x) { return __builtin_sqrt(x); }
}  // namespace srook
