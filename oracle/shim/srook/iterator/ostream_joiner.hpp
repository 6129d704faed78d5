// stand-in (oracle/shim/README.md): output iterator writing its elements separated by a delimiter
#pragma once
#include <iterator>
#include <ostream>
namespace srook {
template <class Delim>
struct ostream_joiner {
    using iterator_category = std::output_iterator_tag;
    using value_type = void;
    using difference_type = void;
    using pointer = void;
    using reference = void;
    std::ostream* os;
    Delim delim;
    bool first = true;
    template <class T>
    ostream_joiner& operator=(const T& v)
    {
        if (!first) *os << delim;
        first = false;
        *os << v;
        return *this;
    }
    ostream_joiner& operator*() { return *this; }
    ostream_joiner& operator++() { return *this; }
    ostream_joiner& operator++(int) { return *this; }
};
template <class Delim>
ostream_joiner<std::decay_t<Delim>> make_ostream_joiner(std::ostream& os, Delim&& d) { return {&os, std::forward<Delim>(d)}; }
}  // namespace srook
