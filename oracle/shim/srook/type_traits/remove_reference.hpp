// stand-in (oracle/shim/README.md)
#pragma once
#include <srook/type_traits/is_same.hpp>
