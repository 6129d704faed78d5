// stand-in (oracle/shim/README.md): for_each over a "counter" range calls f(element, index)
#pragma once
#include <cstddef>
#include <functional>
#include <initializer_list>
#include <vector>
namespace srook {
namespace algorithm {
template <class Range>
struct counter_range {
    Range& r;
};
template <class T>
struct counter_list {
    std::vector<T> items;
};
template <class Range>
counter_range<Range> make_counter(Range& r) { return {r}; }
template <class T>
counter_list<T> make_counter(std::initializer_list<T> il) { return {std::vector<T>(il)}; }
}  // namespace algorithm
template <class Range, class F>
F for_each(algorithm::counter_range<Range> c, F f)
{
    std::size_t i = 0;
    for (auto& v : c.r) f(v, i++);
    return f;
}
template <class T, class F>
F for_each(algorithm::counter_list<T> c, F f)
{
    std::size_t i = 0;
    for (auto& v : c.items) f(static_cast<typename T::type&>(v), i++);   // std::reference_wrapper elements
    return f;
}
}  // namespace srook
