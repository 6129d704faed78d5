// stand-in (oracle/shim/README.md)
#pragma once
#include <srook/config/feature/constexpr.hpp>
