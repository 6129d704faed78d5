// stand-in (oracle/shim/README.md).  A null char pointer gives an empty view: the reference's main() builds views of
// argv[argc] (src/encoder/main.cpp:67-69, src/decoder/main.cpp:99-104), which std::string_view makes undefined.
#pragma once
#include <string_view>
namespace srook {
struct string_view : std::string_view {
    using std::string_view::string_view;
    constexpr string_view(const char* s) : std::string_view(s ? std::string_view(s) : std::string_view()) {}
    constexpr string_view(std::string_view s) : std::string_view(s) {}
};
}  // namespace srook
