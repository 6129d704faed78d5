// stand-in (oracle/shim/README.md)
#pragma once
#include <boost/range/algorithm/copy.hpp>
