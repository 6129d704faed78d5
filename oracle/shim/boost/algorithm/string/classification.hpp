// stand-in (oracle/shim/README.md)
#pragma once
#include <cctype>
namespace boost {
struct is_space_pred {
    bool operator()(char c) const { return std::isspace(static_cast<unsigned char>(c)) != 0; }
};
inline is_space_pred is_space() { return {}; }
}  // namespace boost
