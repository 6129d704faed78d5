// stand-in (oracle/shim/README.md): feature macros of SrookCppLibraries under a C++17 compiler
#pragma once
#include <cassert>
#include <cstddef>
#include <type_traits>
// the real library's headers pull these in transitively; the reference relies on that (e.g. std::unique_ptr in
// src/decoder/jpezy_decoder.hpp:90 without <memory>)
#include <cstdlib>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#define SROOK_CONSTEXPR constexpr
#define SROOK_CONSTEXPR_OR_CONST constexpr
#define SROOK_NOEXCEPT_TRUE noexcept
#define SROOK_NOEXCEPT(...)
#define SROOK_IF_CONSTEXPR if constexpr
#define SROOK_DECLTYPE(...) decltype(__VA_ARGS__)
#define SROOK_DEDUCED_TYPENAME typename
#define SROOK_STRONG_ENUM_BEGIN(name) enum class name
#define SROOK_STRONG_ENUM_EPILOG(name)
#define SROOK_FINAL final
#define SROOK_THROW throw
#define SROOK_TRY try
#define SROOK_CATCH(...) catch (__VA_ARGS__)
#define SROOK_ATTRIBUTE_FALLTHROUGH [[fallthrough]]
#define SROOK_ST_ASSERT(...) static_assert(__VA_ARGS__, "SROOK_ST_ASSERT")
#define SROOK_REQUIRES(...) std::enable_if_t<(__VA_ARGS__), std::nullptr_t> = nullptr
#define REQUIRES(...) SROOK_REQUIRES(__VA_ARGS__)
