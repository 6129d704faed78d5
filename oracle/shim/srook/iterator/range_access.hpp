// stand-in (oracle/shim/README.md)
#pragma once
#include <iterator>
namespace srook {
using std::begin;
using std::end;
}  // namespace srook
